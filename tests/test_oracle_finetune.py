"""CPU: the finetune oracle (oracle/finetune.py) against tests/golden/finetune.npz - the outputs of the REAL
reference's few_shot_style_finetune_losses + backward + AdamW (tests/golden/make_golden_finetune.py) - and the host
logic of the training loop: timestep sampler, batch sharding, the data-parallel gradient rule over gloo (world 2),
the training structs of the C-ABI."""
import ctypes
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "tests", "golden"))

import finetune_inputs as FI  # noqa: E402
from helpers import relerr  # noqa: E402
from oracle import finetune as OF  # noqa: E402
from oracle import schedule as OSch  # noqa: E402
from oracle.weights import NoiseTape, encoder_layer_weights, mdm_state_dict, text_features  # noqa: E402


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(os.path.join(REPO, "tests", "golden", "finetune.npz")))


def _mask(shape):
    from mst_b200.data_loaders.stylexia_posrot_utils import get_inpainting_mask
    return torch.from_numpy(get_inpainting_mask("root_horizontal", shape)).float()


def _oracle_case(cfg):
    inp = FI.make_inputs()
    T = inp["content"].shape[-1]
    front, enc, menc = mdm_state_dict(181, seed=0), FI.style_encoder_state(), FI.motion_encoder_state()
    w_enc = {k: v.clone().requires_grad_(True) for k, v in enc.items()}
    sch = OSch.Schedule(OSch.cosine_betas(1000), OSch.space_timesteps(1000, cfg["respacing"]))
    tape = NoiseTape(17)
    tape.draw(inp["content"].shape)  # the reference's unused th.randn_like(x_content_start)
    terms = OF.finetune_losses(
        sch, front, w_enc, menc, inp["x_start"], inp["t"], inp["content"], inp["style"], text_features(inp["texts_style"]),
        torch.ones(1, T, dtype=torch.bool), _mask((1, 181, 1, T)), tape, text_feat_t2m=text_features(inp["texts_t2m"]),
        frame_mask_t2m=inp["frame_mask_t2m"], inp_mask_t2m=_mask(tuple(inp["x_start"].shape)),
        skip_steps=cfg["skip_steps"], semantic_guidance=cfg["semantic_guidance"], use_ddim=cfg["use_ddim"], Ls=10.0,
        noise_t2m=inp["noise_t2m"])
    terms["loss"].backward()
    return terms, w_enc


@pytest.mark.parametrize("case", ["ddim_sg0", "ddim_sg1", "ddpm_sg0"])
def test_oracle_finetune_matches_reference_golden(gold, case):
    terms, w_enc = _oracle_case(dict(FI.cases())[case])
    assert abs(terms["loss"].item() - float(gold[f"{case}/loss"])) < 1e-5 * abs(float(gold[f"{case}/loss"]))
    assert relerr(terms["rot_mse"].detach(), gold[f"{case}/rot_mse"]) < 1e-5
    if f"{case}/text_cosine" in gold:
        assert abs(terms["text_cosine"].item() - float(gold[f"{case}/text_cosine"])) < 1e-5
    dig = gold[f"{case}/grad_digest"]
    for i, n in enumerate(str(s) for s in gold["param_names"]):
        got = FI.digest(w_enc[n].grad)
        assert abs(got[0] - dig[i][0]) < 1e-4 * abs(dig[i][0]), n
    for key in gold:
        if key.startswith(f"{case}/grad/"):
            assert relerr(w_enc[key.split("/grad/")[1]].grad, gold[key]) < 1e-4, key


def test_oracle_motion_encoder_matches_reference_golden(gold):
    inp = FI.make_inputs()
    x = inp["x_start"].clone().requires_grad_(True)
    mu = OF.motion_encoder_forward(mdm_state_dict(181, seed=0), FI.motion_encoder_state(), x, inp["frame_mask_t2m"])
    assert relerr(mu.detach(), gold["menc/mu"]) < 1e-5
    (mu * torch.from_numpy(gold["menc/w"])).sum().backward()
    assert relerr(x.grad, gold["menc/dx"]) < 1e-4


def test_oracle_adamw_is_torch_adamw():
    g = torch.Generator().manual_seed(0)
    p0 = torch.randn(1000, generator=g)
    p_ref = p0.clone().requires_grad_(True)
    opt = torch.optim.AdamW([p_ref], lr=1e-4, weight_decay=0.01)
    p, m, v = p0.clone(), torch.zeros(1000), torch.zeros(1000)
    for step in range(1, 4):
        grad = torch.randn(1000, generator=g)
        p_ref.grad = grad.clone()
        opt.step()
        OF.adamw_step(p, grad, m, v, step, lr=1e-4, wd=0.01)
        assert torch.allclose(p, p_ref.detach(), rtol=0, atol=1e-7)


def test_training_structs_abi():
    from mst_b200 import _lib as L
    lib = L.load()
    a, b = ctypes.c_size_t(), ctypes.c_size_t()
    assert lib.mst_abi_sizes_train(ctypes.byref(a), ctypes.byref(b)) == 0
    assert a.value == ctypes.sizeof(L.LayerGrads) == 12 * 8
    assert b.value == ctypes.sizeof(L.BackwardArgs)


def test_uniform_sampler_follows_numpy_rng_and_range():
    from mst_b200.diffusion.resample import create_named_schedule_sampler

    class D:
        num_timesteps = 20

    s = create_named_schedule_sampler("uniform", D())
    np.random.seed(3)
    t, w = s.sample(64, "cpu", range(6))
    np.random.seed(3)
    want = np.random.choice(range(6), size=(64,), p=np.ones(6) / 6)
    assert t.tolist() == want.tolist() and t.dtype == torch.int64
    assert torch.allclose(w, torch.ones(64))
    with pytest.raises(NotImplementedError):
        create_named_schedule_sampler("loss-second-moment", D())


def test_shard_batch_slices_per_sample_entries_only():
    from mst_b200.train.training_loop import shard_batch
    with pytest.raises(ValueError):  # ragged split: the equal-weight all-reduce would bias the t2m term
        shard_batch(torch.zeros(5, 3, 1, 2), {"y": {}}, 0, 2)
    with pytest.raises(ValueError):  # fewer samples than ranks: an empty shard would hang the collective
        shard_batch(torch.zeros(1, 3, 1, 2), {"y": {}}, 1, 2)
    B = 6
    batch = torch.arange(B * 6, dtype=torch.float32).view(B, 3, 1, 2)
    cond = {"y": {"text": [f"t{i}" for i in range(B)], "mask": torch.ones(B, 1, 1, 2), "lengths": torch.arange(B),
                  "scalar": 3, "table": torch.zeros(7)}}
    seen = []
    for r in range(2):
        b, c, (lo, hi) = shard_batch(batch, cond, r, 2)
        assert torch.equal(b, batch[lo:hi]) and c["y"]["text"] == cond["y"]["text"][lo:hi]
        assert c["y"]["lengths"].tolist() == list(range(lo, hi)) and c["y"]["table"].shape == (7,)
        seen += list(range(lo, hi))
    assert seen == list(range(B))


# ------------------------------------------------------------------ data-parallel gradient rule, gloo world 2
def _small_states(d=64, ff=128, n_layers=2):
    front = mdm_state_dict(n_feats=24, d=d, ff=ff, n_layers=n_layers, clip_dim=64, seed=0, pe_len=1000)
    enc, menc = {}, {}
    for i in range(n_layers):
        enc.update(encoder_layer_weights(f"seqTransEncoder.layers.{i}.", d, ff, 7))
        menc.update(encoder_layer_weights(f"seqTransEncoder.layers.{i}.", d, ff, 5))
    g = torch.Generator().manual_seed(9)
    menc["muQuery"], menc["sigmaQuery"] = torch.randn(1, d, generator=g), torch.randn(1, d, generator=g)
    return front, enc, menc


def _small_loss(lo, hi, w_enc):
    front, _, menc = _small_states()
    B, F, T = 4, 24, 10
    g = torch.Generator().manual_seed(1)
    x_start, noise_t2m = torch.randn(B, F, 1, T, generator=g), torch.rand(B, F, 1, T, generator=g)
    content, style = torch.randn(1, F, 1, T, generator=g), torch.randn(1, F, 1, T, generator=g)
    feat_t2m, feat_style = torch.randn(B, 64, generator=g), torch.randn(1, 64, generator=g)
    t = torch.tensor([1, 4, 2, 0])
    fm = torch.arange(T)[None, :] < torch.tensor([10, 7, 9, 5])[:, None]
    imask = torch.zeros(1, F, 1, T)
    imask[:, :3] = 1.0
    sch = OSch.Schedule(OSch.cosine_betas(1000), OSch.space_timesteps(1000, "ddim20"))
    tape = NoiseTape(4)
    return OF.finetune_losses(sch, front, w_enc, menc, x_start[lo:hi], t[lo:hi], content, style, feat_style,
                              torch.ones(1, T, dtype=torch.bool), imask, tape, text_feat_t2m=feat_t2m[lo:hi],
                              frame_mask_t2m=fm[lo:hi], inp_mask_t2m=imask.expand(hi - lo, -1, -1, -1), skip_steps=900,
                              semantic_guidance=1, use_ddim=1, Ls=10.0, noise_t2m=noise_t2m[lo:hi])["loss"]


def _dp_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mst_b200.train.training_loop import shard_batch
        _, enc, _ = _small_states()
        w_enc = {k: v.clone().requires_grad_(True) for k, v in enc.items()}
        _, _, (lo, hi) = shard_batch(torch.zeros(4, 1), {"y": {}}, rank, world)
        _small_loss(lo, hi, w_enc).backward()
        flat = torch.cat([w_enc[k].grad.flatten() for k in sorted(w_enc)])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)   # what TrainInpaintingLoop.sync_gradients does on the arena
        ret[rank] = flat / world                      # ... and FusedAdamW applies with grad_scale = 1 / world
    finally:
        dist.destroy_process_group()


def test_data_parallel_gradient_rule_world2_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ret = mp.Manager().dict()
    mp.spawn(_dp_worker, args=(2, port, ret), nprocs=2, join=True)
    _, enc, _ = _small_states()
    w_enc = {k: v.clone().requires_grad_(True) for k, v in enc.items()}
    _small_loss(0, 4, w_enc).backward()
    want = torch.cat([w_enc[k].grad.flatten() for k in sorted(w_enc)])
    assert torch.equal(ret[0], ret[1])
    assert relerr(ret[0], want) < 1e-5   # mean over the full t2m batch == mean of the equal-shard means
