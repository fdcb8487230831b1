"""Batch sharding of the sampling loop across the GPUs of one box (one process per GPU).

The reference has no distributed path (``utils/dist_util.py:18-41`` is a stub that only records a device id).
Every operation of the sampling loop is per-sample, so the path *partitions*: rank r runs the unchanged
``p_sample_loop`` on a contiguous slice of the batch with no collective inside the loop; the in-kernel Philox
noise is keyed by the GLOBAL sample index, so the value computed for sample i does not depend on the number of
ranks.  The only (optional) communication is one all_gather of the finished samples.
"""
from __future__ import annotations

from typing import Optional

import torch


def shard_bounds(n_samples: int, rank: int, world: int):
    """Contiguous, balanced slice [start, start+count) of ``n_samples`` owned by ``rank`` (the first
    ``n_samples % world`` ranks hold one extra sample).  Pure integer arithmetic."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world: {rank}/{world}")
    if n_samples < 0:
        raise ValueError("n_samples must be >= 0")
    base, extra = divmod(n_samples, world)
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


def shard_model_kwargs(model_kwargs: dict, start: int, count: int, n_samples: int) -> dict:
    """Slice every per-sample entry of ``model_kwargs['y']`` (tensors / lists whose leading dimension is the
    batch) to this rank's samples; everything else is passed through."""
    y = model_kwargs.get('y', {})
    out = {}
    for k, v in y.items():
        if isinstance(v, torch.Tensor) and v.dim() >= 1 and v.shape[0] == n_samples:
            out[k] = v[start:start + count].contiguous()
        elif isinstance(v, (list, tuple)) and len(v) == n_samples:
            out[k] = type(v)(v[start:start + count])
        else:
            out[k] = v
    res = dict(model_kwargs)
    res['y'] = out
    return res


def sample_sharded(diffusion, model, shape, model_kwargs, *, rank: Optional[int] = None, world: Optional[int] = None,
                   gather: bool = True, use_ddim: bool = False, noise=None, init_image=None, **loop_kwargs):
    """Run ``diffusion.p_sample_loop`` (or ``ddim_sample_loop``) for the GLOBAL batch ``shape`` with the samples
    split across the ranks of the default process group.

    Returns the full ``[B, ...]`` result on every rank when ``gather`` (one all_gather of the final samples), else
    this rank's ``[count, ...]`` shard.  With ``diffusion.rng == 'philox'`` the result equals the single-process
    result for the same seed, bit for bit."""
    import torch.distributed as dist
    if world is None:
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if world > 1 else 0
    B = int(shape[0])
    start, count = shard_bounds(B, rank, world)
    local_shape = (count,) + tuple(shape[1:])
    kw = shard_model_kwargs(model_kwargs, start, count, B)
    saved = getattr(diffusion, "philox_sample_offset", 0)
    diffusion.philox_sample_offset = saved + start
    try:
        fn = diffusion.ddim_sample_loop if use_ddim else diffusion.p_sample_loop
        sl = slice(start, start + count)
        local = None
        if count > 0:
            local = fn(model, local_shape, noise=None if noise is None else noise[sl].contiguous(),
                       init_image=None if init_image is None else init_image[sl].contiguous(), model_kwargs=kw,
                       **loop_kwargs)
    finally:
        diffusion.philox_sample_offset = saved
    if not gather or world == 1:
        return local
    return gather_shards(local, B, local_shape, world, device=None if local is None else local.device)


def gather_shards(local, n_samples: int, local_shape, world: int, device=None, dtype=torch.float32):
    """all_gather of ragged shards: every rank pads to the largest shard, gathers, and strips the padding."""
    import torch.distributed as dist
    max_count = -(-n_samples // world)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    pad = torch.zeros((max_count,) + tuple(local_shape[1:]), dtype=dtype, device=device)
    if local is not None and local.shape[0] > 0:
        pad[: local.shape[0]].copy_(local)
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    parts = []
    for r in range(world):
        _, c = shard_bounds(n_samples, r, world)
        parts.append(bufs[r][:c])
    return torch.cat(parts, dim=0)
