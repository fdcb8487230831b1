"""CPU, world_size 2 over gloo: the host-side sharding logic of the multi-GPU sampling path
(mst_b200/sharding.py).  The device work is replaced by a fake per-rank sampler whose output is a pure
function of the GLOBAL sample index - exactly the property the Philox-keyed CUDA path has - so the test
checks that shards partition the batch, that per-sample kwargs are sliced consistently, that the Philox
offset handed to the sampler is the global index of the shard's first sample, and that the gathered result
equals the single-process result for ragged and empty shards."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mst_b200 import sharding


def test_shard_bounds_partition():
    for n in (0, 1, 2, 7, 64, 4096, 4097):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            for (s0, c0), (s1, _) in zip(spans, spans[1:]):
                assert s0 + c0 == s1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        sharding.shard_bounds(4, 2, 2)


def test_shard_model_kwargs_slices_only_batch_entries():
    B = 5
    kw = {"y": {"text": [f"t{i}" for i in range(B)], "scale": torch.arange(B).float(), "uncond": False,
                "inpainting_mask": torch.arange(B * 6).float().view(B, 3, 1, 2), "table": torch.zeros(7)}}
    out = sharding.shard_model_kwargs(kw, 1, 3, B)["y"]
    assert out["text"] == ["t1", "t2", "t3"] and out["scale"].tolist() == [1.0, 2.0, 3.0]
    assert out["inpainting_mask"].shape == (3, 3, 1, 2) and out["inpainting_mask"][0, 0, 0, 0] == 6.0
    assert out["uncond"] is False and out["table"].shape == (7,)
    assert kw["y"]["scale"].shape == (B,)  # caller's dict untouched


class _FakeDiffusion:
    """p_sample_loop stand-in: sample value = f(global index, per-sample kwargs)."""
    rng = "philox"
    philox_sample_offset = 0

    def p_sample_loop(self, model, shape, noise=None, init_image=None, model_kwargs=None, **kw):
        n = shape[0]
        idx = torch.arange(n, dtype=torch.float32) + self.philox_sample_offset
        out = idx.view(n, 1, 1, 1).expand(shape).clone()
        out += model_kwargs["y"]["scale"].view(n, 1, 1, 1) * 1000.0
        if noise is not None:
            out += noise
        return out

    ddim_sample_loop = p_sample_loop


def _worker(rank, world, port, B, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        shape = (B, 3, 1, 4)
        kw = {"y": {"scale": torch.arange(B).float() * 2, "text": ["x"] * B}}
        noise = torch.arange(B * 12, dtype=torch.float32).view(shape) * 1e-3
        d = _FakeDiffusion()
        full = sharding.sample_sharded(d, None, shape, kw, noise=noise, clip_denoised=False)
        local = sharding.sample_sharded(d, None, shape, kw, noise=noise, gather=False)
        assert d.philox_sample_offset == 0  # restored
        ret[rank] = (full, local)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [4, 5, 1])
def test_sample_sharded_world2_gloo(B):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    world = 2
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, B, ret), nprocs=world, join=True)
    shape = (B, 3, 1, 4)
    kw = {"y": {"scale": torch.arange(B).float() * 2, "text": ["x"] * B}}
    noise = torch.arange(B * 12, dtype=torch.float32).view(shape) * 1e-3
    want = _FakeDiffusion().p_sample_loop(None, shape, noise=noise, model_kwargs=kw)
    for r in range(world):
        full, local = ret[r]
        assert torch.equal(full, want)  # same on every rank, equal to the single-process result
        s0, c = sharding.shard_bounds(B, r, world)
        if c == 0:
            assert local is None
        else:
            assert torch.equal(local, want[s0:s0 + c])
