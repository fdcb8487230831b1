"""GPU: the finetune path - taped forward, explicit backward, MotionEncoder,
masked-L2 loss, differentiable sampling, AdamW / norms - against (a) the oracle's torch-CPU autograd on the same
seeded inputs and (b) tests/golden/finetune.npz, produced by the REAL reference
(tests/golden/make_golden_finetune.py: StyleDiffusion + few_shot_style_finetune_losses + backward + AdamW).

Tolerance: fp32 everywhere, gradients / losses within 2e-4 relative (max |a-b| / max |b| per tensor); the oracle
itself sits within 1e-6 of the reference.
"""
import os

import numpy as np
import pytest
import torch

from helpers import Args, relerr
from oracle import denoiser as OD
from oracle import finetune as OF
from oracle.weights import NoiseTape, mdm_state_dict, text_features

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 2e-4
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def FI_digest(t):
    t = t.detach().double().flatten().cpu()
    return np.array([t.norm().item(), t.sum().item()] + t[:14].tolist(), dtype=np.float64)


def _golden_module():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_finetune_inputs",
                                                  os.path.join(REPO, "tests", "golden", "finetune_inputs.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(os.path.join(REPO, "tests", "golden", "finetune.npz")))


@pytest.fixture(scope="module")
def FI():
    return _golden_module()


def _style_model(mu, FI, tmp_path_factory, precision="fp32"):
    """mst StyleDiffusion with the golden script's deterministic weights (frozen MDM front, own encoder, MotionEncoder)."""
    from mst_b200.model.mdm_forstyledataset import StyleDiffusion
    tmp = tmp_path_factory.mktemp("ckpt")
    front, enc, menc = mdm_state_dict(181, seed=0), FI.style_encoder_state(), FI.motion_encoder_state()
    torch.save(front, tmp / "mdm.pt")
    torch.save(menc, tmp / "menc.pt")
    args = Args()
    args.mdm_path, args.semantic_discriminator_path = str(tmp / "mdm.pt"), str(tmp / "menc.pt")
    model = StyleDiffusion(**mu.get_transfer_args(args))
    missing, unexpected = model.load_state_dict(enc, strict=False)
    assert not unexpected and all(k.startswith("motion_enc.") for k in missing)
    model.mst_train_precision = precision
    return model.to(DEV).eval(), front, enc, menc


@pytest.fixture(scope="module")
def mu():
    from mst_b200.utils import model_util
    assert torch.cuda.is_available()
    torch.cuda.set_device(0)
    return model_util


@pytest.fixture(scope="module")
def style(mu, FI, tmp_path_factory):
    return _style_model(mu, FI, tmp_path_factory)


def test_taped_forward_equals_inference_forward(style):
    model, *_ = style
    model.mst_precision = "fp32"
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 181, 1, 20, generator=g).to(DEV)
    t = torch.tensor([999, 10], device=DEV)
    y = {"text_feat": text_features(["a", "b"]).to(DEV), "text": ["a", "b"]}
    with torch.no_grad():
        ref = model(x, t, y)
    out = model(x, t, y)  # grad mode: taped fp32 path
    assert out.requires_grad and out.grad_fn is not None
    assert relerr(out.detach(), ref) < 2e-5


# T <= 79 runs the fused attention kernels (five CTAs of 16 query rows per (sequence, head) while B * heads <= 32, one CTA of
# 80 rows above that: B = 12), T = 100 the general batched-GEMM path
@pytest.mark.parametrize("B,T", [(2, 20), (1, 76), (3, 33), (2, 100), (12, 76)])
def test_denoiser_backward_matches_oracle_autograd(style, B, T):
    model, front, enc, _ = style
    g = torch.Generator().manual_seed(100 + T)
    x = torch.randn(B, 181, 1, T, generator=g)
    d_out = torch.randn(B, 181, 1, T, generator=g)
    t = torch.randint(0, 1000, (B,), generator=g)
    feat = text_features((["walk", "run", "jump"] * 4)[:B])
    # oracle: torch-CPU autograd over the plain tensor algebra
    w_enc = {k: v.clone().requires_grad_(True) for k, v in enc.items()}
    xo = x.clone().requires_grad_(True)
    out_o = OD.mdm_forward(front, xo, t, feat, enc_w=w_enc)
    (out_o * d_out).sum().backward()
    # product
    model.zero_grad(set_to_none=True)
    xg = x.to(DEV).requires_grad_(True)
    out = model(xg, t.to(DEV), {"text_feat": feat.to(DEV), "text": ["x"] * B})
    assert relerr(out.detach(), out_o.detach()) < 1e-4
    (out * d_out.to(DEV)).sum().backward()
    assert relerr(xg.grad, xo.grad) < TOL
    worst = 0.0
    for name, p in model.seqTransEncoder.named_parameters():
        e = relerr(p.grad, w_enc["seqTransEncoder." + name].grad)
        worst = max(worst, e)
        assert e < TOL, (name, e)
    print(f"B={B} T={T}: worst parameter-gradient rel err {worst:.2e}")
    # gradients accumulate (torch .grad semantics): a second backward doubles them
    before = model.seqTransEncoder.layers[3].linear1.weight.grad.clone()
    out2 = model(x.to(DEV), t.to(DEV), {"text_feat": feat.to(DEV), "text": ["x"] * B})
    (out2 * d_out.to(DEV)).sum().backward()
    assert relerr(model.seqTransEncoder.layers[3].linear1.weight.grad, 2 * before) < 1e-5


def test_motion_encoder_matches_reference_golden(style, gold, FI):
    model, *_ = style
    inp = FI.make_inputs()
    x = inp["x_start"].to(DEV).requires_grad_(True)
    y = {"mask": inp["frame_mask_t2m"][:, None, None, :].to(DEV), "text_feat": text_features(inp["texts_t2m"]).to(DEV)}
    mu_, feat = model.motion_enc(x, y)
    assert relerr(mu_.detach(), gold["menc/mu"]) < 1e-4
    assert torch.equal(feat.cpu(), text_features(inp["texts_t2m"]))
    (mu_ * torch.from_numpy(gold["menc/w"]).to(DEV)).sum().backward()
    assert relerr(x.grad, gold["menc/dx"]) < TOL


def test_motion_encoder_tensor_core_projections_close_to_fp32(mu, FI, tmp_path_factory):
    """B * T >= 512 routes the MotionEncoder's in-projection (forward) and its backward through the tcgen05 training GEMM in
    the 16-bit mode (csrc/train.cu::inproj_tc / inproj_bwd_tc); mu and d mu / d x stay within the 16-bit tolerance of the fp32
    mode, ragged lengths included."""
    B, T = 8, 76
    g = torch.Generator().manual_seed(5)
    x0 = torch.randn(B, 181, 1, T, generator=g)
    lengths = torch.randint(T // 2, T + 1, (B,), generator=g)
    mask = (torch.arange(T)[None, :] < lengths[:, None])[:, None, None, :].to(DEV)
    wgt = torch.randn(B, 512, generator=g).to(DEV)
    y = {"mask": mask, "text_feat": text_features(["a"] * B).to(DEV)}
    out = {}
    for prec in ("fp32", "bf16"):
        model, *_ = _style_model(mu, FI, tmp_path_factory, precision=prec)
        model.motion_enc.mst_train_precision = prec
        x = x0.clone().to(DEV).requires_grad_(True)
        mu_, _ = model.motion_enc(x, y)
        (mu_ * wgt).sum().backward()
        out[prec] = (mu_.detach().clone(), x.grad.clone())
    assert relerr(out["bf16"][0], out["fp32"][0]) < 2e-2
    assert rel_l2(out["bf16"][1], out["fp32"][1]) < 3e-2


def test_masked_l2_kernel_forward_backward():
    from mst_b200.diffusion.gaussian_diffusion import GaussianDiffusion  # noqa: F401
    from mst_b200 import engine as K
    g = torch.Generator().manual_seed(5)
    R, F, T = 6, 181, 20
    a = torch.randn(1, F, 1, T, generator=g)
    b = torch.randn(R, F, 1, T, generator=g).requires_grad_(True)
    mask = (torch.arange(T) < 13).float().view(1, 1, 1, T)
    want = OF.masked_l2(a.expand(R, -1, -1, -1), b, mask.expand(R, -1, -1, -1))
    gl = torch.randn(R, generator=g)
    (want * gl).sum().backward()
    got = K.masked_l2_forward(a.to(DEV), b.detach().to(DEV), mask.to(DEV))
    assert relerr(got, want.detach()) < 1e-6
    gb = K.masked_l2_backward(a.to(DEV), b.detach().to(DEV), mask.to(DEV), gl.to(DEV))
    assert relerr(gb, b.grad) < 1e-6


def _patched_noise(tape, noise_t2m):
    import torch as th
    orig = th.randn, th.randn_like, th.rand_like

    class _Ctx:
        def __enter__(self):
            th.randn = lambda *s, **k: tape.draw(
                s[0] if len(s) == 1 and isinstance(s[0], (tuple, list, torch.Size)) else s).to(k.get("device", "cpu"))
            th.randn_like = lambda a, **k: tape.draw(a.shape).to(a.device)
            th.rand_like = lambda a, **k: noise_t2m.clone().to(a.device)

        def __exit__(self, *exc):
            th.randn, th.randn_like, th.rand_like = orig
    return _Ctx()


def _finetune_terms(mu, model, FI, cfg):
    from mst_b200.data_loaders.stylexia_posrot_utils import get_inpainting_mask
    inp = FI.make_inputs()
    F, T = 181, inp["content"].shape[-1]
    diffusion = mu.create_gaussian_diffusion(Args(), mu.InpaintingGaussianDiffusion, timestep_respacing=cfg["respacing"])
    mask_style = torch.from_numpy(get_inpainting_mask("root_horizontal", (1, F, 1, T))).float().to(DEV)
    mask_t2m = torch.from_numpy(get_inpainting_mask("root_horizontal", tuple(inp["x_start"].shape))).float().to(DEV)
    style_kwargs = {"y": {"text": inp["texts_style"], "text_feat": text_features(inp["texts_style"]).to(DEV),
                          "mask": torch.ones(1, 1, 1, T, dtype=torch.bool, device=DEV), "lengths": torch.tensor([T]),
                          "inpainted_motion": inp["style"].to(DEV), "inpainting_mask": mask_style}}
    t2m_kwargs = {"y": {"text": inp["texts_t2m"], "text_feat": text_features(inp["texts_t2m"]).to(DEV),
                        "mask": inp["frame_mask_t2m"][:, None, None, :].to(DEV), "lengths": torch.tensor(inp["lengths"]),
                        "inpainting_mask": mask_t2m}}
    with _patched_noise(NoiseTape(17), inp["noise_t2m"]):
        terms = diffusion.few_shot_style_finetune_losses(
            model, inp["x_start"].to(DEV), inp["t"].to(DEV), inp["content"].to(DEV), inp["style"].to(DEV),
            skip_steps=cfg["skip_steps"], model_kwargs=style_kwargs, model_t2m_kwargs=t2m_kwargs,
            semantic_guidance=cfg["semantic_guidance"], use_ddim=cfg["use_ddim"], Ls=10)
    return terms


@pytest.mark.parametrize("case", ["ddim_sg0", "ddim_sg1", "ddpm_sg0"])
def test_finetune_losses_and_gradients_match_reference(mu, FI, gold, tmp_path_factory, case):
    model, *_ = _style_model(mu, FI, tmp_path_factory)
    cfg = dict(FI.cases())[case]
    terms = _finetune_terms(mu, model, FI, cfg)
    assert abs(terms["loss"].item() - float(gold[f"{case}/loss"])) < 1e-4 * abs(float(gold[f"{case}/loss"]))
    assert relerr(terms["rot_mse"].detach(), gold[f"{case}/rot_mse"]) < 1e-4
    if cfg["semantic_guidance"]:
        assert abs(terms["text_cosine"].item() - float(gold[f"{case}/text_cosine"])) < 1e-4
    model.zero_grad(set_to_none=True)
    terms["loss"].backward()
    names = [str(n) for n in gold["param_names"]]
    params = dict(model.named_parameters())
    trainable = [n for n, p in model.named_parameters() if p.requires_grad]
    assert trainable == names  # the same 96 tensors as the reference, in the same order
    dig = gold[f"{case}/grad_digest"]
    worst = 0.0
    for i, n in enumerate(names):
        got = FI_digest(params[n].grad)
        scale = max(abs(dig[i][0]) / np.sqrt(params[n].numel()), 1e-12)  # rms of the reference gradient
        assert abs(got[0] - dig[i][0]) < TOL * abs(dig[i][0]), (n, got[0], dig[i][0])
        assert np.abs(got[2:] - dig[i][2:]).max() < 5 * TOL * max(scale, np.abs(dig[i][2:]).max()), n
        worst = max(worst, abs(got[0] - dig[i][0]) / abs(dig[i][0]))
    print(f"{case}: worst gradient-norm rel err over 96 tensors {worst:.2e}")
    for key in gold:
        if key.startswith(f"{case}/grad/"):
            assert relerr(params[key.split("/grad/")[1]].grad, gold[key]) < TOL, key
    assert relerr(params["seqTransEncoder.layers.0.linear1.weight"].grad[:8, :64],
                  gold[f"{case}/grad_slice/layers.0.linear1.weight"]) < TOL
    assert relerr(params["seqTransEncoder.layers.7.self_attn.in_proj_weight"].grad[510:518, :64],
                  gold[f"{case}/grad_slice/layers.7.self_attn.in_proj_weight"]) < TOL


def test_trainer_norms_adamw_step_match_reference(mu, FI, gold, tmp_path_factory):
    """zero_grad -> loss -> backward -> norms -> fused AdamW (training_loop.py:196-200) == torch.optim.AdamW of the
    reference run, tensor by tensor."""
    from mst_b200.diffusion.fp16_util import MixedPrecisionTrainer
    from mst_b200.train.training_loop import FusedAdamW
    case = "ddim_sg1"
    model, *_ = _style_model(mu, FI, tmp_path_factory)
    trainer = MixedPrecisionTrainer(model=model)
    opt = FusedAdamW(trainer.flat, lr=1e-4, weight_decay=0.0, model=model)
    trainer.zero_grad()
    terms = _finetune_terms(mu, model, FI, dict(FI.cases())[case])
    trainer.backward(terms["loss"])
    grad_norm, param_norm = trainer._compute_norms()
    dig = gold[f"{case}/grad_digest"]
    assert abs(grad_norm - np.sqrt((dig[:, 0] ** 2).sum())) < TOL * grad_norm
    want_pn = np.sqrt(sum(float(p.detach().double().pow(2).sum()) for p in model.parameters()))
    assert abs(param_norm - want_pn) < 1e-5 * want_pn
    trainer.optimize(opt)
    after = gold[f"{case}/param_after_digest"]
    params = dict(model.named_parameters())
    for i, n in enumerate(str(s) for s in gold["param_names"]):
        got = FI_digest(params[n])
        # Adam's first step moves every element by lr * g / (|g| + eps) ~ +-1e-4: an element whose gradient is ~1e-8
        # may legitimately land anywhere in +-lr; everything else must agree to fp32 rounding
        assert abs(got[0] - after[i][0]) < 1e-5 * abs(after[i][0]), n
        diff = np.abs(got[2:] - after[i][2:])
        assert diff.max() < 2.1e-4 and (diff < 2e-6).sum() >= 12, (n, diff)
    # the optimiser wrote through raw pointers: the engines must notice (weights re-packed on next use)
    with torch.no_grad():
        y = {"text_feat": text_features(["a"]).to(DEV), "text": ["a"]}
        x = torch.zeros(1, 181, 1, 20, device=DEV)
        model.mst_precision = "fp32"
        out_new = model(x, torch.tensor([5], device=DEV), y)
    w_enc = {k: v.detach().cpu() for k, v in model.state_dict().items() if k.startswith("seqTransEncoder.")}
    want = OD.mdm_forward(mdm_state_dict(181, seed=0), x.cpu(), torch.tensor([5]), text_features(["a"]), enc_w=w_enc)
    assert relerr(out_new, want) < 1e-4


def test_training_loop_runs_and_reduces_the_loss(mu, FI, tmp_path_factory):
    from mst_b200.data_loaders.stylexia_posrot_utils import get_inpainting_mask
    from mst_b200.train.training_loop import TrainInpaintingLoop
    model, *_ = _style_model(mu, FI, tmp_path_factory)
    inp = FI.make_inputs()
    F, T = 181, inp["content"].shape[-1]

    class A(Args):
        batch_size, lr, weight_decay, lr_anneal_steps, style_finetune, semantic_guidance = 3, 1e-4, 0.0, 0, 1, 1
        skip_steps, use_ddim, Ls, num_steps = 700, 1, 10, 4

    diffusion = mu.create_gaussian_diffusion(A(), mu.InpaintingGaussianDiffusion, timestep_respacing="ddim20")
    mask_style = torch.from_numpy(get_inpainting_mask("root_horizontal", (1, F, 1, T))).float().to(DEV)
    mask_t2m = torch.from_numpy(get_inpainting_mask("root_horizontal", tuple(inp["x_start"].shape))).float().to(DEV)
    style_cond = {"y": {"text": inp["texts_style"], "text_feat": text_features(inp["texts_style"]).to(DEV),
                        "mask": torch.ones(1, 1, 1, T, dtype=torch.bool, device=DEV), "lengths": torch.tensor([T]),
                        "inpainted_motion": inp["style"].to(DEV), "inpainting_mask": mask_style}}
    cond = {"y": {"text": inp["texts_t2m"], "text_feat": text_features(inp["texts_t2m"]).to(DEV),
                  "mask": inp["frame_mask_t2m"][:, None, None, :].to(DEV), "lengths": torch.tensor(inp["lengths"]),
                  "inpainting_mask": mask_t2m}}
    loop = TrainInpaintingLoop(A(), None, model, [(inp["x_start"], cond)], diffusion=diffusion,
                               style_data=((inp["content"].to(DEV), style_cond),))
    np.random.seed(0)
    torch.manual_seed(0)
    losses = []
    for _ in range(4):
        loop.run_step(inp["x_start"].to(DEV), cond, inp["content"].to(DEV), style_cond)
        losses.append(float(loop.last_losses["loss"]))
    assert all(np.isfinite(losses))
    assert losses[-1] < losses[0], losses
    assert loop.step == 4 and loop.opt.step_count == 4


def test_side_stream_step_equals_the_serial_step(mu, FI, tmp_path_factory, monkeypatch):
    """The trainer runs the text-to-motion branch on its own stream and back-propagates it into a second gradient arena
    (training_loop.py::_early_t2m_backward); MST_OVERLAP_ALLREDUCE=0 is the serial step.  Same seeds, dropout off:
    the parameters after three steps and the logged losses agree to fp32 rounding, and the concurrent mode really used
    the side arena."""
    from mst_b200.data_loaders.stylexia_posrot_utils import get_inpainting_mask
    from mst_b200.train.training_loop import TrainInpaintingLoop
    inp = FI.make_inputs()
    F, T = 181, inp["content"].shape[-1]

    class A(Args):
        batch_size, lr, weight_decay, lr_anneal_steps, style_finetune, semantic_guidance = 3, 1e-4, 0.0, 0, 1, 1
        skip_steps, use_ddim, Ls, num_steps = 700, 1, 10, 4

    mask_style = torch.from_numpy(get_inpainting_mask("root_horizontal", (1, F, 1, T))).float().to(DEV)
    mask_t2m = torch.from_numpy(get_inpainting_mask("root_horizontal", tuple(inp["x_start"].shape))).float().to(DEV)
    style_cond = {"y": {"text": inp["texts_style"], "text_feat": text_features(inp["texts_style"]).to(DEV),
                        "mask": torch.ones(1, 1, 1, T, dtype=torch.bool, device=DEV), "lengths": torch.tensor([T]),
                        "inpainted_motion": inp["style"].to(DEV), "inpainting_mask": mask_style}}
    cond = {"y": {"text": inp["texts_t2m"], "text_feat": text_features(inp["texts_t2m"]).to(DEV),
                  "mask": inp["frame_mask_t2m"][:, None, None, :].to(DEV), "lengths": torch.tensor(inp["lengths"]),
                  "inpainting_mask": mask_t2m}}
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("MST_OVERLAP_ALLREDUCE", mode)
        model, *_ = _style_model(mu, FI, tmp_path_factory)
        for m in model.modules():
            if hasattr(m, "mst_train_dropout"):
                m.mst_train_dropout = 0.0
        diffusion = mu.create_gaussian_diffusion(A(), mu.InpaintingGaussianDiffusion, timestep_respacing="ddim20")
        loop = TrainInpaintingLoop(A(), None, model, [(inp["x_start"], cond)], diffusion=diffusion,
                                   style_data=((inp["content"].to(DEV), style_cond),))
        np.random.seed(0)
        torch.manual_seed(0)
        torch.cuda.manual_seed(0)
        losses = []
        for _ in range(3):
            loop.run_step(inp["x_start"].to(DEV), cond, inp["content"].to(DEV), style_cond)
            losses.append(float(loop.last_losses["loss"]))
        torch.cuda.synchronize()
        res[mode] = (loop.mp_trainer.flat.train_params.clone(), losses, loop.__dict__.get("_g_t2m"),
                     loop.mp_trainer.last_norms)
    a, b = res["1"][0], res["0"][0]
    # Adam moves an element by ~lr * g / |g|: where the gradient is rounding noise (|g| ~ 1e-9) the two summation orders
    # may step in different directions; everywhere else the parameters agree to fp32 rounding
    diff = (a - b).abs()
    print("side-stream vs serial: max |dp| %.3e, fraction above 2e-6: %.3e" % (float(diff.max()), float((diff > 2e-6).float().mean())))
    assert float(diff.max()) < 1e-4 and float((diff > 2e-6).float().mean()) < 1e-4
    assert np.allclose(res["1"][1], res["0"][1], rtol=1e-5)
    assert np.allclose(res["1"][3], res["0"][3], rtol=1e-4)
    assert res["0"][2] is None and res["1"][2] is not None and float(res["1"][2].abs().sum()) > 0


def test_run_loop_checkpoints_in_the_reference_format_and_resumes(mu, FI, tmp_path_factory):
    """run_loop (reference train/training_loop.py:143-190, :309-348): num_steps // len(data) + 1 epochs, CPU-side
    kwargs moved to the device, modelNNNNNNNNN.pt without the frozen motion_enc. / clip_model. entries, optNNNNNNNNN.pt
    in torch.optim.AdamW's own layout over list(model.parameters()), and a second loop that resumes from them."""
    import os
    from mst_b200.data_loaders.stylexia_posrot_utils import get_inpainting_mask
    from mst_b200.train.training_loop import TrainInpaintingLoop
    save_dir = str(tmp_path_factory.mktemp("ckpt"))
    inp = FI.make_inputs()
    F, T = 181, inp["content"].shape[-1]

    class A(Args):
        batch_size, lr, weight_decay, lr_anneal_steps, style_finetune, semantic_guidance = 3, 1e-4, 0.0, 0, 1, 0
        skip_steps, use_ddim, Ls, num_steps = 700, 1, 10, 2
        log_interval, save_interval, resume_checkpoint, overwrite = 1, 2, "", True

    A.save_dir = save_dir

    def make_loop(args):
        model, *_ = _style_model(mu, FI, tmp_path_factory)
        diffusion = mu.create_gaussian_diffusion(args, mu.InpaintingGaussianDiffusion, timestep_respacing="ddim20")
        # everything on the HOST, as a data loader delivers it: run_loop moves it (reference :155, :166-167)
        style_cond = {"y": {"text": inp["texts_style"], "text_feat": text_features(inp["texts_style"]),
                            "mask": torch.ones(1, 1, 1, T, dtype=torch.bool), "lengths": torch.tensor([T]),
                            "inpainted_motion": inp["style"].clone(),
                            "inpainting_mask": torch.from_numpy(get_inpainting_mask("root_horizontal", (1, F, 1, T))).float()}}
        cond = {"y": {"text": inp["texts_t2m"], "text_feat": text_features(inp["texts_t2m"]),
                      "mask": inp["frame_mask_t2m"][:, None, None, :].clone(), "lengths": torch.tensor(inp["lengths"]),
                      "inpainting_mask": torch.from_numpy(get_inpainting_mask("root_horizontal", tuple(inp["x_start"].shape))).float()}}
        data = [(inp["x_start"].clone(), cond)]  # len(data) == 1 -> num_epochs = num_steps + 1 = 3
        return model, TrainInpaintingLoop(args, None, model, data, diffusion=diffusion, style_data=[(inp["content"].clone(), style_cond)])

    np.random.seed(0)
    model, loop = make_loop(A())
    assert loop.num_epochs == 3
    loop.run_loop()
    assert loop.step == 3
    files = sorted(os.listdir(save_dir))
    # saved after loop indices 0 and 2 (step % save_interval == 0 before the increment); the final save is skipped
    # because (step - 1) % save_interval == 0 - exactly the reference's bookkeeping (:183-190)
    assert files == ["model000000000.pt", "model000000002.pt", "opt000000000.pt", "opt000000002.pt"], files
    sd = torch.load(os.path.join(save_dir, "model000000002.pt"), map_location="cpu")
    assert not any(k.startswith(("motion_enc.", "clip_model.")) for k in sd)
    trainable = {n for n, p in model.named_parameters() if p.requires_grad}
    assert trainable <= set(sd) and len(trainable) == 96
    for n, p in model.named_parameters():
        if n in sd:
            assert torch.equal(sd[n], p.detach().cpu()), n
    # the optimizer file loads into a stock torch.optim.AdamW over list(model.parameters()), as the reference resumes it
    osd = torch.load(os.path.join(save_dir, "opt000000002.pt"), map_location="cpu")
    params = list(model.parameters())
    stock = torch.optim.AdamW(params, lr=1e-4, weight_decay=0.0)
    stock.load_state_dict(osd)
    idx = {id(p): i for i, p in enumerate(params)}
    some = next(p for p in params if p.requires_grad)
    assert torch.equal(stock.state[some]["exp_avg"].cpu(), osd["state"][idx[id(some)]]["exp_avg"])
    assert int(stock.state[some]["step"]) == 3 and len(osd["state"]) == 96
    # resume: a fresh loop pointed at the directory picks up the latest files
    B = A()
    B.resume_checkpoint = save_dir
    model2, loop2 = make_loop(B)
    assert loop2.resume_step == 2 and loop2.opt.step_count == 3
    for (n, p), (_, q) in zip(model.named_parameters(), model2.named_parameters()):
        if p.requires_grad:
            assert torch.equal(p, q), n
    assert torch.equal(loop2.opt.exp_avg, loop.opt.exp_avg)
    loop2.run_loop(max_steps=1)
    assert loop2.step == 1 and "model000000002.pt" in os.listdir(save_dir) and loop2.ckpt_file_name() == "model000000003.pt"


# ------------------------------------------------------------------ bf16 tensor-core training mode
def rel_l2(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.mark.parametrize("B,T", [(2, 20), (1, 76), (5, 60), (12, 76)])
def test_bf16_backward_close_to_fp32_oracle(mu, FI, tmp_path_factory, B, T):
    """MST_TRAIN_PRECISION=bf16: the linear layers (forward, dX, dW) run on the tcgen05 kernel with bf16 operands and
    fp32 accumulation.  Tolerance 3e-2 relative L2 per gradient tensor against the fp32 oracle (BASELINE's bf16 bar is
    2e-2 on x0 predictions; gradients pass through 8 more bf16 roundings on the way back)."""
    model, front, enc, _ = _style_model(mu, FI, tmp_path_factory, precision="bf16")
    g = torch.Generator().manual_seed(200 + T)
    x = torch.randn(B, 181, 1, T, generator=g)
    d_out = torch.randn(B, 181, 1, T, generator=g)
    t = torch.randint(0, 1000, (B,), generator=g)
    feat = text_features([f"c{i}" for i in range(B)])
    w_enc = {k: v.clone().requires_grad_(True) for k, v in enc.items()}
    xo = x.clone().requires_grad_(True)
    out_o = OD.mdm_forward(front, xo, t, feat, enc_w=w_enc)
    (out_o * d_out).sum().backward()
    xg = x.to(DEV).requires_grad_(True)
    out = model(xg, t.to(DEV), {"text_feat": feat.to(DEV), "text": ["x"] * B})
    assert rel_l2(out.detach(), out_o.detach()) < 2e-2
    (out * d_out.to(DEV)).sum().backward()
    assert rel_l2(xg.grad, xo.grad) < 3e-2
    worst = 0.0
    for name, p in model.seqTransEncoder.named_parameters():
        e = rel_l2(p.grad, w_enc["seqTransEncoder." + name].grad)
        worst = max(worst, e)
        assert e < 3e-2, (name, e)
    print(f"bf16 B={B} T={T}: worst parameter-gradient rel-L2 err {worst:.2e}")


def test_bf16_finetune_loss_close_to_reference(mu, FI, gold, tmp_path_factory):
    model, *_ = _style_model(mu, FI, tmp_path_factory, precision="bf16")
    case = "ddim_sg1"
    terms = _finetune_terms(mu, model, FI, dict(FI.cases())[case])
    assert abs(terms["loss"].item() - float(gold[f"{case}/loss"])) < 2e-2 * abs(float(gold[f"{case}/loss"]))
    model.zero_grad(set_to_none=True)
    terms["loss"].backward()
    dig = gold[f"{case}/grad_digest"]
    params = dict(model.named_parameters())
    got = np.sqrt(sum(float(params[str(n)].grad.double().pow(2).sum()) for n in gold["param_names"]))
    want = np.sqrt((dig[:, 0] ** 2).sum())
    assert abs(got - want) < 3e-2 * want


# ------------------------------------------------------------------ dropout of the training forward
def test_dropout_scale_statistics():
    from mst_b200 import engine as K
    n, p = 1 << 20, 0.1
    seed = torch.tensor([12345], dtype=torch.int64, device=DEV)
    m = K.dropout_scale(n, p, seed, 17)
    vals = torch.unique(m)
    assert vals.numel() == 2 and vals[0] == 0 and abs(float(vals[1]) - 1 / 0.9) < 1e-6
    assert abs(float((m == 0).float().mean()) - p) < 2e-3
    assert abs(float(m.mean()) - 1.0) < 3e-3                       # inverted dropout keeps the expectation
    assert not torch.equal(m, K.dropout_scale(n, p, seed, 18))      # another site: another mask
    assert not torch.equal(m, K.dropout_scale(n, p, seed + 1, 17))  # another key: another mask
    assert torch.equal(m, K.dropout_scale(n, p, seed, 17))          # counter-based: reproducible


def _masked_oracle_forward(front, w, x, t, feat, masks, n_heads=4):
    """oracle.denoiser.mdm_forward with the dropout multipliers of the CUDA path applied at the same five sites
    (token sequence after PE; per layer: attention probabilities, out-proj output, GELU output, linear2 output)."""
    import math
    B, F, _, T = x.shape
    S = T + 1
    emb = OD.time_embedding(front, t) + feat @ front["embed_text.weight"].T + front["embed_text.bias"]
    xs = x.permute(3, 0, 1, 2).reshape(T, B, F) @ front["input_process.poseEmbedding.weight"].T + \
        front["input_process.poseEmbedding.bias"]
    seq = torch.cat([emb[None], xs], dim=0) + front["sequence_pos_encoder.pe"][:S]

    def tok(m, width):  # token-major [B, S, width] multipliers -> [S, B, width]
        return m.view(B, S, width).permute(1, 0, 2)

    d = seq.shape[-1]
    dh = d // n_heads
    seq = seq * tok(masks[0], d)
    for l in range(8):
        pre = f"seqTransEncoder.layers.{l}."
        qkv = seq @ w[pre + "self_attn.in_proj_weight"].T + w[pre + "self_attn.in_proj_bias"]
        q, k, v = (z.reshape(S, B, n_heads, dh).permute(1, 2, 0, 3) for z in qkv.split(d, dim=-1))
        att = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(dh), dim=-1) * masks[8 * (l + 1) + 1].view(B, n_heads, S, S)
        o = (att @ v).permute(2, 0, 1, 3).reshape(S, B, d)
        sa = (o @ w[pre + "self_attn.out_proj.weight"].T + w[pre + "self_attn.out_proj.bias"]) * tok(masks[8 * (l + 1) + 2], d)
        seq = OD.layer_norm(seq + sa, w[pre + "norm1.weight"], w[pre + "norm1.bias"])
        h = OD.gelu(seq @ w[pre + "linear1.weight"].T + w[pre + "linear1.bias"]) * tok(masks[8 * (l + 1) + 3], 1024)
        ff = (h @ w[pre + "linear2.weight"].T + w[pre + "linear2.bias"]) * tok(masks[8 * (l + 1) + 4], d)
        seq = OD.layer_norm(seq + ff, w[pre + "norm2.weight"], w[pre + "norm2.bias"])
    out = seq[1:] @ front["output_process.poseFinal.weight"].T + front["output_process.poseFinal.bias"]
    return out.reshape(T, B, F, 1).permute(1, 2, 3, 0).contiguous()


@pytest.mark.parametrize("pooled", [False, True])
def test_dropout_forward_backward_matches_masked_oracle(mu, FI, tmp_path_factory, pooled):
    """p = 0.1 dropout (the reference's train-mode setting): the masks are counter-based, so the test reads them back
    through the test hook, applies them in the oracle's tensor algebra, and compares outputs and all 96 gradients."""
    from mst_b200 import engine as K
    model, front, enc, _ = _style_model(mu, FI, tmp_path_factory, precision="fp32")
    model.mst_train_dropout = 0.1
    B, T, p = 2, 19, 0.1
    S, M = T + 1, B * (T + 1)
    g = torch.Generator().manual_seed(77)
    x = torch.randn(B, 181, 1, T, generator=g)
    d_out = torch.randn(B, 181, 1, T, generator=g)
    t = torch.tensor([400, 20])
    feat = text_features(["a", "b"])
    if pooled:  # the trainer's mode: pooled tape, CUDA-graph replay, gradients accumulated straight into .grad
        from mst_b200.diffusion.fp16_util import MixedPrecisionTrainer
        trainer = MixedPrecisionTrainer(model=model)
        trainer.zero_grad()
    else:
        model.zero_grad(set_to_none=True)
    y = {"text_feat": feat.to(DEV), "text": ["x"] * B}
    for rep_ in range(2 if pooled else 1):  # second pooled pass replays the captured graphs with a NEW key
        if pooled:
            trainer.zero_grad()
        torch.manual_seed(1000 + rep_)
        xg = x.to(DEV).requires_grad_(not pooled)
        out = model(xg, t.to(DEV), y)
        (out * d_out.to(DEV)).sum().backward()
        torch.manual_seed(1000 + rep_)
        key = int(torch.randint(0, 2 ** 62, (1,)).item())
        sizes = {0: S * 512}   # per sequence: every sequence of a call has its own key (key + index in the call)
        for l in range(8):
            sizes.update({8 * (l + 1) + 1: 4 * S * S, 8 * (l + 1) + 2: S * 512, 8 * (l + 1) + 3: S * 1024,
                          8 * (l + 1) + 4: S * 512})
        seeds = [torch.tensor([key + b], dtype=torch.int64, device=DEV) for b in range(B)]
        masks = {site: torch.cat([K.dropout_scale(n, p, seeds[b], site) for b in range(B)]).cpu()
                 for site, n in sizes.items()}
        w_enc = {k: v.clone().requires_grad_(True) for k, v in enc.items()}
        xo = x.clone().requires_grad_(True)
        out_o = _masked_oracle_forward(front, w_enc, xo, t, feat, masks)
        (out_o * d_out).sum().backward()
        assert relerr(out.detach(), out_o.detach()) < 1e-4
        if not pooled:
            assert relerr(xg.grad, xo.grad) < TOL
        for name, prm in model.seqTransEncoder.named_parameters():
            assert relerr(prm.grad, w_enc["seqTransEncoder." + name].grad) < TOL, (name, rep_)
    # and the masks matter: the deterministic output differs
    model.mst_train_dropout = 0.0
    with torch.no_grad():
        assert relerr(model(x.to(DEV), t.to(DEV), y), out_o.detach()) > 1e-2


def test_batched_backward_of_sequential_forwards(mu, FI, tmp_path_factory):
    """Trainer mode: forwards of single sequences recorded one by one on a shared tape (the six differentiable DDIM
    steps) are back-propagated by ONE batched launch sequence once the last of them reports its output gradient -
    the accumulated gradients equal the sum of the individual plain backward passes, with and without dropout."""
    from mst_b200.diffusion.fp16_util import MixedPrecisionTrainer
    T, n_calls = 76, 5
    g = torch.Generator().manual_seed(31)
    xs = [torch.randn(1, 181, 1, T, generator=g) for _ in range(n_calls)]
    ds = [torch.randn(1, 181, 1, T, generator=g) for _ in range(n_calls)]
    ts = [torch.randint(0, 1000, (1,), generator=g) for _ in range(n_calls)]
    feat = text_features(["style"])
    y = {"text_feat": feat.to(DEV), "text": ["x"]}
    for p_drop in (0.0, 0.1):
        # plain: fresh tape per forward, gradients returned to autograd
        plain, *_ = _style_model(mu, FI, tmp_path_factory, precision="fp32")
        plain.mst_train_dropout = p_drop
        plain.zero_grad(set_to_none=True)
        torch.manual_seed(5)
        outs_plain = []
        for x, d_, t in zip(xs, ds, ts):
            out = plain(x.to(DEV), t.to(DEV), y)
            outs_plain.append(out.detach().clone())
            (out * d_.to(DEV)).sum().backward()
        # pooled: shared tape, deferred batched backward (some gradients staged out of order)
        pooled, *_ = _style_model(mu, FI, tmp_path_factory, precision="fp32")
        pooled.mst_train_dropout = p_drop
        trainer = MixedPrecisionTrainer(model=pooled)
        for rep_ in range(2):   # second round replays the captured graphs
            trainer.zero_grad()
            torch.manual_seed(5)
            outs = [pooled(x.to(DEV), t.to(DEV), y) for x, t in zip(xs, ts)]
            for o, ref_o in zip(outs, outs_plain):
                assert relerr(o.detach(), ref_o) < 1e-6
            order = [2, 0, 4, 1, 3]
            slot = next(iter(pooled._mst_tape_slots.values()))[0][0]
            for i, k in enumerate(order):
                (outs[k] * ds[k].to(DEV)).sum().backward()
                assert (len(slot.pending) == i + 1) if i + 1 < n_calls else (len(slot.pending) == 0)
            for (name, a), (_, b) in zip(pooled.seqTransEncoder.named_parameters(), plain.seqTransEncoder.named_parameters()):
                assert relerr(a.grad, b.grad) < 2e-5, (name, p_drop, rep_)
        # a forward whose output never reaches the loss must not block the others: the trainer flushes
        trainer.zero_grad()
        torch.manual_seed(5)
        outs = [pooled(x.to(DEV), t.to(DEV), y) for x, t in zip(xs[:3], ts[:3])]
        loss = (outs[0] * ds[0].to(DEV)).sum() + (outs[2] * ds[2].to(DEV)).sum()
        trainer.backward(loss)
        assert float(pooled.seqTransEncoder.layers[0].linear1.weight.grad.abs().sum()) > 0


def test_backward_humanml_feature_width(mu):
    """F = 263 (humanml) and T = 196: the taped forward / backward of an MDM whose front end is frozen (only the feature
    width of the in / out projections changes; S = 197 takes the general attention path)."""
    from mst_b200.model.mdm_forstyledataset import MDM

    class A(Args):
        dataset = "humanml"

    state = mdm_state_dict(263, seed=3)
    model = MDM(**mu.get_transfer_args(A()))
    model.load_state_dict(state, strict=False)
    model.to(DEV).eval()
    for name, p in model.named_parameters():
        p.requires_grad_(name.startswith("seqTransEncoder."))
    model.mst_train_precision = "fp32"
    B, T = 2, 196
    g = torch.Generator().manual_seed(9)
    x, d_out = torch.randn(B, 263, 1, T, generator=g), torch.randn(B, 263, 1, T, generator=g)
    t = torch.tensor([10, 900])
    feat = text_features(["a", "b"])
    w_enc = {k: v.clone().requires_grad_(True) for k, v in state.items() if k.startswith("seqTransEncoder.")}
    xo = x.clone().requires_grad_(True)
    out_o = OD.mdm_forward(state, xo, t, feat, enc_w=w_enc)
    (out_o * d_out).sum().backward()
    xg = x.to(DEV).requires_grad_(True)
    out = model(xg, t.to(DEV), {"text_feat": feat.to(DEV), "text": ["x"] * B})
    assert relerr(out.detach(), out_o.detach()) < 1e-4
    (out * d_out.to(DEV)).sum().backward()
    assert relerr(xg.grad, xo.grad) < TOL
    for name, p in model.named_parameters():
        if p.requires_grad:
            assert relerr(p.grad, w_enc[name].grad) < TOL, name
