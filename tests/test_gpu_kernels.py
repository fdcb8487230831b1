"""GPU: every kernel behind the C-ABI against the CPU oracle (or a torch fp32 restatement for the two
bf16 tensor-core kernels) on the same seeded inputs.

Tolerances (written where they are used):
  * integer / index work (timestep maps, masks, the device timestep counter)  -> bit-exact
  * fused update / q_sample / cfg lerp (fp32 elementwise)                     -> <= 2 ulp-level, rel 1e-6
  * fp32 denoiser forward                                                    -> rel 1e-4  (BASELINE north_star)
  * bf16 tcgen05 kernels                                                     -> rel 2e-2  (BASELINE north_star)
"""
import math

import numpy as np
import pytest
import torch

from helpers import relerr
from oracle import denoiser as OD
from oracle import philox as OP
from oracle import sampler as OS
from oracle import schedule as OSch
from oracle.weights import NoiseTape, mdm_state_dict

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


@pytest.fixture(scope="module")
def K():
    from mst_b200 import engine
    assert torch.cuda.is_available(), "these tests need a B200"
    torch.cuda.set_device(0)
    return engine


@pytest.fixture(scope="module")
def L():
    from mst_b200 import _lib
    return _lib


@pytest.fixture(scope="module")
def state():
    return mdm_state_dict(n_feats=181, seed=0)


def _tabs(sch, eta=0.0):
    f = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).float().to(DEV)
    ab, abp = sch.abar, sch.abar_prev
    sig = eta * np.sqrt((1 - abp) / (1 - ab)) * np.sqrt(1 - ab / abp)
    return dict(c1=f(sch.coef1), c2=f(sch.coef2), sigma=f(np.exp(0.5 * sch.logvar)), sqrt_ab=f(sch.sqrt_abar),
                sqrt_1m_ab=f(sch.sqrt_1m_abar), recip=f(sch.sqrt_recip_abar), recipm1=f(sch.sqrt_recipm1_abar),
                ddim_c1=f(np.sqrt(abp)), ddim_c2=f(np.sqrt(1 - abp - sig ** 2)), ddim_sigma=f(sig))


def _rand(shape, seed):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed))


def _mask(shape, rows=(0, 1, 2)):
    m = torch.zeros(shape)
    m[:, list(rows)] = 1.0
    return m


# ----------------------------------------------------------------------------- fused update
@pytest.mark.parametrize("shape", [(3, 181, 1, 196), (2, 181, 1, 76), (2, 263, 1, 61), (1, 5, 1, 3)])
@pytest.mark.parametrize("clip", [False, True])
def test_update_ddpm_matches_oracle(K, L, shape, clip):
    sch = OSch.Schedule(OSch.cosine_betas(1000), OSch.space_timesteps(1000, "50"))
    tabs = _tabs(sch)
    oc, ou, x, inp, eps = (_rand(shape, s) for s in (1, 2, 3, 4, 5))
    mask = _mask(shape)
    scale = torch.tensor([2.5, 0.0, 1.0][: shape[0]])
    t = torch.tensor([49, 0, 17][: shape[0]])
    model_out = ou + scale.view(-1, 1, 1, 1) * (oc - ou)
    want, want_x0 = OS.p_sample(sch, model_out, x, t, eps, mask, inp, clip)
    got, got_x0 = torch.empty(shape, device=DEV), torch.empty(shape, device=DEV)
    K.update_step(sampler=L.SAMPLER_DDPM, out_cond=oc.to(DEV), out_uncond=ou.to(DEV), cfg_scale=scale.to(DEV),
                  x_t=x.to(DEV), x_prev=got, pred_xstart=got_x0, mask=mask.to(DEV), x_inpaint=inp.to(DEV),
                  clip_denoised=clip, t_vec=t.to(DEV), coef1=tabs["c1"], coef2=tabs["c2"], sigma=tabs["sigma"],
                  noise_kind=L.NOISE_TENSOR, noise=eps.to(DEV))
    # x0 is a pure fp32 blend in the reference's operation order: bit-exact
    assert torch.equal(got_x0.cpu(), want_x0)
    # the sample differs only through sigma = exp(0.5*logvar): fp64-then-round here, fp32 exp in the reference
    assert relerr(got, want) < 1e-6
    nz = (t == 0).nonzero().flatten().tolist()
    for b in nz:  # t == 0 rows carry no noise: posterior mean only, bit-exact
        assert torch.equal(got[b].cpu(), want[b])


@pytest.mark.parametrize("mask_layout", ["full", "ft", "f", "none"])
def test_update_mask_layouts_agree(K, L, mask_layout):
    shape = (4, 181, 1, 76)
    sch = OSch.Schedule(OSch.cosine_betas(1000))
    tabs = _tabs(sch)
    oc, x, inp, eps = (_rand(shape, s).to(DEV) for s in (1, 3, 4, 5))
    full = _mask(shape, rows=(0, 1, 2, 30)).to(DEV)
    mask = {"full": full, "ft": full[0].reshape(181, 76).contiguous(), "f": full[0, :, 0, 0].contiguous(), "none": None}[mask_layout]
    t = torch.full((4,), 500, device=DEV)
    ref, got = torch.empty(shape, device=DEV), torch.empty(shape, device=DEV)
    kw = dict(sampler=L.SAMPLER_DDPM, out_cond=oc, x_t=x, clip_denoised=False, t_vec=t, coef1=tabs["c1"], coef2=tabs["c2"],
              sigma=tabs["sigma"], noise_kind=L.NOISE_TENSOR, noise=eps)
    if mask_layout == "none":
        K.update_step(x_prev=got, **kw)
        want, _ = OS.p_sample(sch, oc.cpu(), x.cpu(), t.cpu(), eps.cpu())
        assert relerr(got, want) < 1e-6
        return
    K.update_step(x_prev=ref, mask=full, x_inpaint=inp, **kw)
    K.update_step(x_prev=got, mask=mask, x_inpaint=inp, **kw)
    assert torch.equal(ref, got)  # the compact layouts are the same arithmetic


@pytest.mark.parametrize("eta", [0.0, 0.5])
def test_update_ddim_matches_oracle(K, L, eta):
    shape = (2, 181, 1, 76)
    sch = OSch.Schedule(OSch.cosine_betas(1000), OSch.space_timesteps(1000, "ddim20"))
    tabs = _tabs(sch, eta)
    oc, x, inp, eps = (_rand(shape, s) for s in (1, 3, 4, 5))
    mask = _mask(shape)
    t = torch.tensor([5, 0])
    want, want_x0 = OS.ddim_sample(sch, oc, x, t, eps, mask, inp, False, eta=eta)
    got, got_x0 = torch.empty(shape, device=DEV), torch.empty(shape, device=DEV)
    K.update_step(sampler=L.SAMPLER_DDIM, out_cond=oc.to(DEV), x_t=x.to(DEV), x_prev=got, pred_xstart=got_x0,
                  mask=mask.to(DEV), x_inpaint=inp.to(DEV), t_vec=t.to(DEV), coef1=tabs["ddim_c1"], coef2=tabs["ddim_c2"],
                  sigma=tabs["ddim_sigma"], recip=tabs["recip"], recipm1=tabs["recipm1"], noise_kind=L.NOISE_TENSOR,
                  noise=eps.to(DEV))
    assert torch.equal(got_x0.cpu(), want_x0)
    # coefficient tables are rounded from float64 here, recomputed in fp32 by the reference: a few ulp
    assert relerr(got, want) < 2e-6


def test_update_device_timestep_counter(K, L):
    """t lives in device memory and the kernel decrements it: index work, bit-exact."""
    shape = (2, 181, 1, 20)
    sch = OSch.Schedule(OSch.cosine_betas(1000), OSch.space_timesteps(1000, "10"))
    tabs = _tabs(sch)
    oc, eps = _rand(shape, 1).to(DEV), _rand(shape, 5).to(DEV)
    x = _rand(shape, 3).to(DEV)
    x_ref = x.clone()
    t_dev = torch.tensor([9], dtype=torch.int32, device=DEV)
    counter = torch.zeros(1, dtype=torch.int32, device=DEV)
    for step in range(10):
        t_now = 9 - step
        tmp = torch.empty_like(x_ref)
        K.update_step(sampler=L.SAMPLER_DDPM, out_cond=oc, x_t=x_ref, x_prev=tmp, t_imm=t_now, coef1=tabs["c1"],
                      coef2=tabs["c2"], sigma=tabs["sigma"], noise_kind=L.NOISE_TENSOR, noise=eps)
        x_ref = tmp
        K.update_step(sampler=L.SAMPLER_DDPM, out_cond=oc, x_t=x, x_prev=x, t_scalar_dev=t_dev, advance_t=True,
                      block_counter=counter, coef1=tabs["c1"], coef2=tabs["c2"], sigma=tabs["sigma"],
                      noise_kind=L.NOISE_TENSOR, noise=eps)
        assert int(t_dev.item()) == t_now - 1 and int(counter.item()) == 0
        assert torch.equal(x, x_ref)


def test_const_noise_repeats_sample_zero(K, L):
    shape = (3, 181, 1, 20)
    sch = OSch.Schedule(OSch.cosine_betas(1000))
    tabs = _tabs(sch)
    oc, x, eps = _rand(shape, 1).to(DEV), _rand(shape, 3).to(DEV), _rand(shape, 5).to(DEV)
    a, b = torch.empty_like(x), torch.empty_like(x)
    kw = dict(sampler=L.SAMPLER_DDPM, out_cond=oc, x_t=x, t_imm=300, coef1=tabs["c1"], coef2=tabs["c2"],
              sigma=tabs["sigma"], noise_kind=L.NOISE_TENSOR)
    K.update_step(x_prev=a, noise=eps, const_noise=True, **kw)
    K.update_step(x_prev=b, noise=eps[[0]].repeat(3, 1, 1, 1).contiguous(), **kw)
    assert torch.equal(a, b)


def test_q_sample_and_cfg_combine(K):
    shape = (3, 181, 1, 76)
    sch = OSch.Schedule(OSch.cosine_betas(1000))
    tabs = _tabs(sch)
    x0, eps = _rand(shape, 1), _rand(shape, 2)
    mask = _mask(shape)
    t = torch.tensor([999, 0, 400])
    want = OS.q_sample(sch, x0, t, eps, mask)
    got = K.q_sample(x0.to(DEV), eps.to(DEV), mask.to(DEV), t.to(DEV), 0, tabs["sqrt_ab"], tabs["sqrt_1m_ab"])
    assert torch.equal(got.cpu(), want)
    want = OS.q_sample(sch, x0, t, eps, None)
    got = K.q_sample(x0.to(DEV), eps.to(DEV), None, t.to(DEV), 0, tabs["sqrt_ab"], tabs["sqrt_1m_ab"])
    assert torch.equal(got.cpu(), want)
    scale = torch.tensor([2.5, 0.0, -1.0])
    got = K.cfg_combine(x0.to(DEV), eps.to(DEV), scale.to(DEV))
    assert torch.equal(got.cpu(), eps + scale.view(-1, 1, 1, 1) * (x0 - eps))


def test_philox_stream_matches_oracle(K, L):
    B, per = 3, 181 * 76
    got = K.philox_normal((B, 181, 1, 76), seed=0x1234ABCD5678, sample_offset=1 << 33, t=17, device=DEV).cpu().view(B, per)
    want = torch.from_numpy(OP.philox_normal(B, per, 0x1234ABCD5678, 1 << 33, 17))
    # the integer stream is exact; logf / sincospif differ from numpy's by a few ulp
    assert float((got - want).abs().max()) < 2e-5
    assert abs(float(got.mean())) < 0.02 and abs(float(got.std()) - 1.0) < 0.02
    # in-kernel generation == the standalone generator, and shards reproduce the full batch
    shape = (4, 181, 1, 76)
    sch = OSch.Schedule(OSch.cosine_betas(1000))
    tabs = _tabs(sch)
    oc, x = _rand(shape, 1).to(DEV), _rand(shape, 3).to(DEV)
    kw = dict(sampler=L.SAMPLER_DDPM, t_imm=123, coef1=tabs["c1"], coef2=tabs["c2"], sigma=tabs["sigma"], philox_seed=77)
    full, via_tensor = torch.empty_like(x), torch.empty_like(x)
    K.update_step(out_cond=oc, x_t=x, x_prev=full, noise_kind=L.NOISE_PHILOX, philox_sample_offset=10, **kw)
    eps = K.philox_normal(shape, seed=77, sample_offset=10, t=123, device=DEV)
    K.update_step(out_cond=oc, x_t=x, x_prev=via_tensor, noise_kind=L.NOISE_TENSOR, noise=eps, **kw)
    assert torch.equal(full, via_tensor)
    half = torch.empty_like(x[2:])
    K.update_step(out_cond=oc[2:].contiguous(), x_t=x[2:].contiguous(), x_prev=half, noise_kind=L.NOISE_PHILOX,
                  philox_sample_offset=12, **kw)
    assert torch.equal(half, full[2:])


# ----------------------------------------------------------------------------- denoiser pieces
def _engine(K, state, precision, n_feats=181):
    eng = K.Engine(n_feats=n_feats, precision=precision, device=DEV)
    g = lambda k: state[k].to(DEV)
    top = {"in_w": g("input_process.poseEmbedding.weight"), "in_b": g("input_process.poseEmbedding.bias"),
           "pe": g("sequence_pos_encoder.pe").reshape(5000, 512),
           "t_w1": g("embed_timestep.time_embed.0.weight"), "t_b1": g("embed_timestep.time_embed.0.bias"),
           "t_w2": g("embed_timestep.time_embed.2.weight"), "t_b2": g("embed_timestep.time_embed.2.bias"),
           "txt_w": g("embed_text.weight"), "txt_b": g("embed_text.bias"),
           "out_w": g("output_process.poseFinal.weight"), "out_b": g("output_process.poseFinal.bias")}
    layers = []
    for i in range(8):
        p = f"seqTransEncoder.layers.{i}."
        layers.append({"qkv_w": g(p + "self_attn.in_proj_weight"), "qkv_b": g(p + "self_attn.in_proj_bias"),
                       "o_w": g(p + "self_attn.out_proj.weight"), "o_b": g(p + "self_attn.out_proj.bias"),
                       "w1": g(p + "linear1.weight"), "b1": g(p + "linear1.bias"), "w2": g(p + "linear2.weight"),
                       "b2": g(p + "linear2.bias"), "ln1_g": g(p + "norm1.weight"), "ln1_b": g(p + "norm1.bias"),
                       "ln2_g": g(p + "norm2.weight"), "ln2_b": g(p + "norm2.bias")})
    eng.load_weights(top, layers)
    return eng


def test_time_and_text_embedding(K, state):
    eng = _engine(K, state, "fp32")
    t = torch.tensor([0, 1, 500, 999, 4999])
    assert relerr(eng.time_embed(t.to(DEV)), OD.time_embedding(state, t)) < 1e-5
    feat = _rand((5, 512), 7)
    want = feat @ state["embed_text.weight"].T + state["embed_text.bias"]
    assert relerr(eng.text_embed(feat.to(DEV)), want) < 1e-5


@pytest.mark.parametrize("m,n,k", [(128, 256, 64), (300, 256, 512), (25216, 1536, 512), (1000, 512, 1024), (394, 192, 512),
                                   (77, 64, 64)])
def test_tc_gemm_bf16(L, m, n, k):
    import ctypes as C
    lib = L.load()
    a = _rand((m, k), 1).to(DEV).bfloat16().contiguous()
    w = (_rand((n, k), 2) / math.sqrt(k)).to(DEV).bfloat16().contiguous()
    bias = _rand((n,), 3).to(DEV)
    c = torch.full((m, n), float("nan"), device=DEV)
    L.check(lib.mst_test_gemm_bf16(a.data_ptr(), w.data_ptr(), bias.data_ptr(), c.data_ptr(), m, n, k,
                                   torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    want = a.float() @ w.float().T + bias
    # same bf16 inputs, fp32 accumulation on both sides: only summation order differs
    assert relerr(c, want) < 1e-4


@pytest.mark.parametrize("epi,m,n,k", [(0, 25216, 1536, 512), (0, 197, 256, 64), (1, 25216, 1024, 512), (1, 300, 256, 512),
                                       (2, 25216, 512, 512), (2, 25216, 512, 1024), (2, 197, 512, 512), (2, 50, 512, 64),
                                       (2, 128 * 149, 512, 512)])
def test_tc_gemm_fused_epilogues(L, epi, m, n, k):
    """bias / bias+GELU (TMA-store epilogue) and the 2-CTA-cluster residual+LayerNorm GEMM against torch fp32 on the
    same bf16 inputs.  Output is bf16: tolerance = one bf16 rounding (2^-8 relative) plus accumulation order."""
    lib = L.load()
    a = _rand((m, k), 1).to(DEV).bfloat16().contiguous()
    w = (_rand((n, k), 2) / math.sqrt(k)).to(DEV).bfloat16().contiguous()
    bias = _rand((n,), 3).to(DEV)
    res = _rand((m, n), 4).to(DEV).bfloat16().contiguous()
    g, b = (1 + 0.1 * _rand((n,), 5)).to(DEV), (0.1 * _rand((n,), 6)).to(DEV)
    out = torch.full((m, n), float("nan"), device=DEV, dtype=torch.bfloat16)
    L.check(lib.mst_test_gemm_epi_bf16(epi, a.data_ptr(), w.data_ptr(), bias.data_ptr(), res.data_ptr(), g.data_ptr(),
                                       b.data_ptr(), out.data_ptr(), m, n, k, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    want = a.float() @ w.float().T + bias
    if epi == 1:
        want = torch.nn.functional.gelu(want)
    elif epi == 2:
        want = torch.nn.functional.layer_norm(want + res.float(), (n,), g, b, eps=1e-5)
    assert torch.isfinite(out.float()).all()
    assert relerr(out.float(), want) < 6e-3


def test_tc_attention_longest_sequence_and_limit(K, L, state):
    """S = 208 tokens is the longest sequence the tcgen05 attention holds in one TMEM slot; 209 must fail loudly."""
    lib = L.load()
    eng = _engine(K, state, "bf16")
    qkv = _rand((2 * 209, 1536), 12).to(DEV).bfloat16().contiguous()
    out = torch.empty(2 * 209, 512, device=DEV, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="208"):
        L.check(lib.mst_test_attention_bf16(eng._h, qkv.data_ptr(), out.data_ptr(), 2, 209, None, 0,
                                            torch.cuda.current_stream().cuda_stream))


@pytest.mark.parametrize("n_seqs,S", [(1, 197), (5, 197), (3, 77), (2, 61), (4, 21), (2, 128), (2, 129), (3, 208), (2, 1)])
def test_tc_attention_bf16(K, L, state, n_seqs, S):
    lib = L.load()
    eng = _engine(K, state, "bf16")
    d, H, dh = 512, 4, 128
    qkv = _rand((n_seqs * S, 3 * d), 11).to(DEV).bfloat16().contiguous()
    out = torch.full((n_seqs * S, d), float("nan"), device=DEV, dtype=torch.bfloat16)
    L.check(lib.mst_test_attention_bf16(eng._h, qkv.data_ptr(), out.data_ptr(), n_seqs, S, None, 0,
                                        torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    q, k, v = qkv.float().view(n_seqs, S, 3, H, dh).permute(2, 0, 3, 1, 4)
    att = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(dh), dim=-1)
    want = (att @ v).permute(0, 2, 1, 3).reshape(n_seqs * S, d)
    # P is rounded to bf16 before the PV product and the output is bf16: 2e-2 (bf16 mode tolerance)
    assert relerr(out.float(), want) < 2e-2


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 2e-2)])
def test_denoiser_forward_matches_golden(K, state, golden, precision, tol):
    eng = _engine(K, state, precision)
    x, t, feat = (torch.from_numpy(golden[k]) for k in ("fwd_x", "fwd_t", "fwd_feat"))
    temb = eng.time_embed(t.to(DEV))
    text = eng.text_embed(feat.to(DEV))
    got_c = eng.forward(x.to(DEV), temb, text, cfg=False)
    assert relerr(got_c, golden["fwd_out_cond"]) < tol
    got_u = eng.forward(x.to(DEV), temb, None, cfg=False, uncond=True)
    assert relerr(got_u, golden["fwd_out_uncond"]) < tol
    oc, ou = eng.forward(x.to(DEV), temb, text, cfg=True)
    assert relerr(oc, golden["fwd_out_cond"]) < tol and relerr(ou, golden["fwd_out_uncond"]) < tol
    if precision == "fp32":
        # batching the two passes changes nothing but the fp32 summation order (the row count selects between the
        # tiled and the one-warp-per-output SIMT GEMM)
        assert relerr(oc, got_c) < 2e-6 and relerr(ou, got_u) < 2e-6


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 2e-2)])
@pytest.mark.parametrize("B,T,F", [(3, 196, 181), (2, 76, 181), (2, 60, 263)])
def test_denoiser_forward_matches_oracle_shapes(K, precision, tol, B, T, F):
    st = mdm_state_dict(n_feats=F, seed=1)
    eng = _engine(K, st, precision, n_feats=F)
    x, feat = _rand((B, F, 1, T), 5), _rand((B, 512), 6)
    t = torch.tensor([999, 0, 421][:B])
    want_c = OD.mdm_forward(st, x, t, feat)
    want_u = OD.mdm_forward(st, x, t, feat, uncond=True)
    oc, ou = eng.forward(x.to(DEV), eng.time_embed(t.to(DEV)), eng.text_embed(feat.to(DEV)), cfg=True)
    assert relerr(oc, want_c) < tol and relerr(ou, want_u) < tol


# ------------------------------------------------------------------ post-sampling decode (scope row N2)
@pytest.mark.parametrize("name,F,T,J", [("stylexia", 181, 76, 20), ("humanml", 263, 196, 22), ("bandai", 190, 60, 21)])
def test_decode_motion_matches_reference_recover_from_ric(name, F, T, J):
    """inv_transform + recover_from_ric fused into one kernel against tests/golden/decode.npz (the REAL reference's
    recover_from_ric, tests/golden/make_golden_decode.py).  fp32; the three prefix sums run in the reference's order, so
    only cos/sin (device vs host libm) differ: 2e-5 of the largest coordinate."""
    import os
    import numpy as np
    from mst_b200.data_loaders.humanml.scripts.motion_process import decode_motion, recover_from_ric
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "decode.npz"))
    g = torch.Generator().manual_seed(0)   # same draws as make_golden_decode.inputs
    sample = torch.randn(3, F, 1, T, generator=g)
    mean = torch.randn(F, generator=g) * 0.3
    std = torch.rand(F, generator=g) * 0.5 + 0.05
    want = torch.from_numpy(gold[f"{name}/joints"])
    got = decode_motion(sample.to(DEV), mean, std, J)
    assert tuple(got.shape) == (3, 1, T, J, 3)
    assert relerr(got, want) < 2e-5
    # the reference-shaped entry point: de-normalised [B, 1, T, F] in, [B, 1, T, J, 3] out
    denorm = (sample.permute(0, 2, 3, 1) * std + mean).float().to(DEV)
    assert relerr(recover_from_ric(denorm, J), want) < 2e-5
    with pytest.raises(RuntimeError, match="CUDA"):
        recover_from_ric(denorm.cpu(), J)
