"""Batch collation into ``model_kwargs`` (reference ``data_loaders/tensors.py``): the boundary
object ``{'y': {mask, lengths, text, ...}}`` the sampler receives."""
import torch


def lengths_to_mask(lengths, max_len):
    """[B] lengths -> [B, max_len] float mask of valid frames."""
    steps = torch.arange(max_len, device=lengths.device).expand(len(lengths), max_len)
    return (steps < lengths.unsqueeze(1)).float()


def collate_tensors(batch):
    """zero-pad a list of same-rank tensors to their per-dim maximum and stack"""
    dims = batch[0].dim()
    size = (len(batch),) + tuple(max(b.size(i) for b in batch) for i in range(dims))
    canvas = batch[0].new_zeros(size=size)
    for i, b in enumerate(batch):
        view = canvas[i]
        for d in range(dims):
            view = view.narrow(d, 0, b.size(d))
        view.add_(b)
    return canvas


_PASS_THROUGH = ('text', 'tokens', 'file_name', 'action_text', 'style')


def collate(batch):
    """list of {'inp': [J, feats, T], 'lengths', 'text', ...} -> (motion [B,J,feats,T], {'y': {...}})"""
    items = [b for b in batch if b is not None]
    motion = collate_tensors([b['inp'] for b in items])
    if 'lengths' in items[0]:
        lengths = [b['lengths'] for b in items]
    else:
        lengths = [len(b['inp'][0][0]) for b in items]
    lengths = torch.as_tensor(lengths)
    mask = lengths_to_mask(lengths, motion.shape[-1]).unsqueeze(1).unsqueeze(1)
    cond = {'y': {'mask': mask, 'lengths': lengths}}
    for key in _PASS_THROUGH:
        if key in items[0]:
            cond['y'][key] = [b[key] for b in items]
    if 'action' in items[0]:
        cond['y']['action'] = torch.as_tensor([b['action'] for b in items]).unsqueeze(1)
    if 'sty_x' in items[0]:
        sty = collate_tensors([b['sty_x'] for b in items])
        sty_len = torch.as_tensor([b['sty_lengths'] for b in items])
        cond['sty_x'] = sty
        cond['sty_y'] = {'mask': lengths_to_mask(sty_len, sty.shape[-1]).unsqueeze(1).unsqueeze(1), 'lengths': sty_len}
    return motion, cond
