"""Batched evaluation caller (scope row N4): ``CompMDMGeneratedDataset`` of the reference
(``data_loaders/humanml/motion_loaders/comp_v6_model_dataset.py:146-263``), the one in-repo consumer that sweeps a whole
data loader through ``diffusion.p_sample_loop`` with a guidance scale.

Same constructor, attributes (``generated_motion``, ``mm_generated_motion``, ``w_vectorizer``) and ``__getitem__`` /
``__len__`` as the reference; what changes underneath:

* the ``mm_num_repeats`` repetitions of a multimodality batch are ONE sampling call of ``repeats x B`` motions (the
  reference loops over them; with the Philox stream keyed by the global sample index each repetition still gets its own
  noise), so the B200 sees the large batch it needs (0.53 of peak at B=64, 0.60 at B>=4096);
* every sampling call goes through the fused trajectory (one CUDA graph per shape) and the samples come back to the
  host once per batch, not once per sample;
* with ``torch.distributed`` initialised, batches are dealt round-robin to the ranks and the lists are gathered with
  ``all_gather_object`` at the end (no collective inside the sampling loop).
"""
import numpy as np
import torch
from torch.utils.data import Dataset


class CompMDMGeneratedDataset(Dataset):

    def __init__(self, model, diffusion, dataloader, mm_num_samples, mm_num_repeats, max_motion_length, num_samples_limit,
                 scale=1.):
        self.dataloader = dataloader
        self.dataset = dataloader.dataset
        assert mm_num_samples < len(dataloader.dataset)
        use_ddim = False  # hard-coded in the reference (:152)
        clip_denoised = False  # hard-coded in the reference (:153)
        self.max_motion_length = max_motion_length
        sample_fn = diffusion.p_sample_loop if not use_ddim else diffusion.ddim_sample_loop

        real_num_batches = len(dataloader)
        if num_samples_limit is not None:
            real_num_batches = num_samples_limit // dataloader.batch_size + 1
        if mm_num_samples > 0:
            mm_idxs = np.random.choice(real_num_batches, mm_num_samples // dataloader.batch_size + 1, replace=False)
            mm_idxs = np.sort(mm_idxs)
        else:
            mm_idxs = []
        self.mm_idxs = list(int(i) for i in mm_idxs)

        dist = torch.distributed
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        rank = dist.get_rank() if world > 1 else 0
        try:
            device = next(model.parameters()).device
        except (StopIteration, AttributeError):
            device = next(model.model.parameters()).device

        model.eval()
        generated, mm_generated = {}, {}
        with torch.no_grad():
            n_done = 0
            for i, (motion, model_kwargs) in enumerate(dataloader):
                if num_samples_limit is not None and n_done >= num_samples_limit:
                    break
                B = motion.shape[0]
                n_done += B
                if i % world != rank:
                    continue
                y = {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in model_kwargs['y'].items()}
                tokens = [t.split('_') for t in y['tokens']]
                if scale != 1.:
                    y['scale'] = torch.ones(B, device=device) * scale  # reference :183-185
                is_mm = i in self.mm_idxs
                repeat_times = mm_num_repeats if is_mm else 1
                # all repetitions in ONE call: sample r * B + b is repetition r of caption b
                y_rep = {k: (v.repeat((repeat_times,) + (1,) * (v.dim() - 1)) if torch.is_tensor(v) and v.dim() > 0
                             and v.shape[0] == B else (list(v) * repeat_times if isinstance(v, (list, tuple)) and len(v) == B
                                                       else v)) for k, v in y.items()}
                shape = (repeat_times * B,) + tuple(motion.shape[1:])
                sample = sample_fn(model, shape, clip_denoised=clip_denoised, model_kwargs={'y': y_rep}, skip_timesteps=0,
                                   init_image=None, progress=False, dump_steps=None, noise=None, const_noise=False)
                # [R*B, J, 1, T] -> host [R, B, T, J] once per batch
                host = sample.reshape(repeat_times, B, sample.shape[1], -1).permute(0, 1, 3, 2).cpu().numpy()
                lengths = y['lengths'].cpu().numpy()
                generated[i] = [{'motion': host[0, b], 'length': lengths[b], 'caption': y['text'][b], 'tokens': tokens[b],
                                 'cap_len': len(tokens[b])} for b in range(B)]
                if is_mm:
                    mm_generated[i] = [{'caption': y['text'][b], 'tokens': tokens[b], 'cap_len': len(tokens[b]),
                                        'mm_motions': [{'motion': host[r, b], 'length': lengths[b]}
                                                       for r in range(repeat_times)]} for b in range(B)]
        if world > 1:
            parts, mm_parts = [None] * world, [None] * world
            dist.all_gather_object(parts, generated)
            dist.all_gather_object(mm_parts, mm_generated)
            generated = {k: v for part in parts for k, v in part.items()}
            mm_generated = {k: v for part in mm_parts for k, v in part.items()}
        self.generated_motion = [d for i in sorted(generated) for d in generated[i]]
        self.mm_generated_motion = [d for i in sorted(mm_generated) for d in mm_generated[i]]
        self.w_vectorizer = getattr(dataloader.dataset, 'w_vectorizer', None)

    def __len__(self):
        return len(self.generated_motion)

    def __getitem__(self, item):
        data = self.generated_motion[item]
        motion, m_length, caption, tokens = data['motion'], data['length'], data['caption'], data['tokens']
        sent_len = data['cap_len']
        if self.dataset.mode == 'eval':
            normed_motion = motion
            denormed_motion = self.dataset.t2m_dataset.inv_transform(normed_motion)
            renormed_motion = (denormed_motion - self.dataset.mean_for_eval) / self.dataset.std_for_eval
            motion = renormed_motion  # T2M evaluators expect their own normalisation (reference :246-251)
        pos_one_hots = []
        word_embeddings = []
        for token in tokens:
            word_emb, pos_oh = self.w_vectorizer[token]
            pos_one_hots.append(pos_oh[None, :])
            word_embeddings.append(word_emb[None, :])
        pos_one_hots = np.concatenate(pos_one_hots, axis=0)
        word_embeddings = np.concatenate(word_embeddings, axis=0)
        return word_embeddings, pos_one_hots, caption, sent_len, motion, m_length, '_'.join(tokens)
