#!/usr/bin/env python
"""Generate tests/golden/finetune.npz from the REAL reference and check the oracle against it.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_finetune.py

What runs: the unmodified reference ``StyleDiffusion`` (random-init per oracle.weights, its two checkpoint files
written to a temp dir as SURVEY section 8(c) describes), ``creat_ddpm_ddim_diffusion`` and
``InpaintingGaussianDiffusion.few_shot_style_finetune_losses`` followed by ``loss.backward()`` and one
``torch.optim.AdamW`` step - i.e. train/training_loop.py:223-303 without the data loader - in eval mode (dropout
off), with torch.randn / randn_like / rand_like replaced by oracle.weights.NoiseTape draws.  The oracle
(oracle/finetune.py) is asserted to agree before anything is written.
"""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)

from make_golden import Args, install_shims, relerr  # noqa: E402
from oracle import finetune as OF  # noqa: E402
from oracle import schedule as OSch  # noqa: E402
from oracle.weights import NoiseTape, mdm_state_dict, text_features  # noqa: E402

OUT = os.path.join(REPO, "tests", "golden")
from finetune_inputs import cases, digest, make_inputs, motion_encoder_state, style_encoder_state  # noqa: E402


def main():
    torch.manual_seed(0)
    install_shims()
    import torch as th
    from utils import model_util as ref_mu
    from model.mdm_forstyledataset import MotionEncoder, StyleDiffusion
    from data_loaders import stylexia_posrot_utils as ref_masks

    tmp = tempfile.mkdtemp(prefix="mst_golden_")
    front = mdm_state_dict(n_feats=181, seed=0)
    enc = style_encoder_state()
    menc = motion_encoder_state()
    torch.save(front, os.path.join(tmp, "mdm.pt"))
    torch.save(menc, os.path.join(tmp, "menc.pt"))
    args = Args()
    args.mdm_path = os.path.join(tmp, "mdm.pt")
    args.semantic_discriminator_path = os.path.join(tmp, "menc.pt")

    inp = make_inputs()
    F, T = 181, inp["content"].shape[-1]
    arrays = {}
    for name, cfg in cases():
        model, diffusion, _ = ref_mu.creat_ddpm_ddim_diffusion(args, ModelClass=StyleDiffusion,
                                                               timestep_respacing=cfg["respacing"])
        missing, unexpected = model.load_state_dict(enc, strict=False)
        assert not unexpected and all(k.startswith("motion_enc.") for k in missing), (missing[:3], unexpected[:3])
        model.eval()
        trainable = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
        assert len(trainable) == 96 and all(n.startswith("seqTransEncoder.") for n, _ in trainable)

        shape_t2m = tuple(inp["x_start"].shape)
        mask_style = torch.from_numpy(ref_masks.get_inpainting_mask("root_horizontal", (1, F, 1, T))).float()
        mask_t2m = torch.from_numpy(ref_masks.get_inpainting_mask("root_horizontal", shape_t2m)).float()
        style_kwargs = {"y": {"text": inp["texts_style"], "mask": torch.ones(1, 1, 1, T, dtype=torch.bool),
                              "lengths": torch.tensor([T]), "inpainted_motion": inp["style"],
                              "inpainting_mask": mask_style}}
        t2m_kwargs = {"y": {"text": inp["texts_t2m"], "mask": inp["frame_mask_t2m"][:, None, None, :],
                            "lengths": torch.tensor(inp["lengths"]), "inpainting_mask": mask_t2m}}
        tape = NoiseTape(17)
        orig = th.randn, th.randn_like, th.rand_like
        th.randn = lambda *s, **k: tape.draw(s[0] if len(s) == 1 and isinstance(s[0], (tuple, list, torch.Size)) else s)
        th.randn_like = lambda a, **k: tape.draw(a.shape)
        th.rand_like = lambda a, **k: inp["noise_t2m"].clone()
        try:
            terms = diffusion.few_shot_style_finetune_losses(
                model, inp["x_start"], inp["t"], inp["content"], inp["style"], skip_steps=cfg["skip_steps"],
                model_kwargs=style_kwargs, model_t2m_kwargs=t2m_kwargs, semantic_guidance=cfg["semantic_guidance"],
                use_ddim=cfg["use_ddim"], Ls=10)
        finally:
            th.randn, th.randn_like, th.rand_like = orig
        model.zero_grad()
        terms["loss"].backward()
        ref_grads = {n: p.grad.detach().clone() for n, p in trainable}

        # ---- the oracle on the same inputs
        w_enc = {k: v.clone().requires_grad_(True) for k, v in enc.items()}
        sch = OSch.Schedule(OSch.cosine_betas(1000), OSch.space_timesteps(1000, cfg["respacing"]))
        otape = NoiseTape(17)
        otape.draw(inp["content"].shape)  # the reference's unused `noise = th.randn_like(x_content_start)` (:1330)
        ot = OF.finetune_losses(
            sch, front, w_enc, menc, inp["x_start"], inp["t"], inp["content"], inp["style"],
            text_features(inp["texts_style"]), torch.ones(1, T, dtype=torch.bool), mask_style, otape,
            text_feat_t2m=text_features(inp["texts_t2m"]), frame_mask_t2m=inp["frame_mask_t2m"], inp_mask_t2m=mask_t2m,
            skip_steps=cfg["skip_steps"], semantic_guidance=cfg["semantic_guidance"], use_ddim=cfg["use_ddim"], Ls=10.0,
            noise_t2m=inp["noise_t2m"])
        ot["loss"].backward()
        e_loss = abs(ot["loss"].item() - terms["loss"].item()) / abs(terms["loss"].item())
        e_grad = max(relerr(w_enc[n].grad, g) for n, g in ref_grads.items())
        print(f"[{name}] loss ref {terms['loss'].item():.6f} oracle {ot['loss'].item():.6f} (rel {e_loss:.1e}); "
              f"max grad rel err over 96 tensors {e_grad:.2e}")
        assert e_loss < 1e-5 and e_grad < 2e-4, (e_loss, e_grad)

        # ---- one AdamW step (train/training_loop.py:97-99: lr 1e-4, weight_decay 0)
        opt = torch.optim.AdamW([p for _, p in trainable], lr=1e-4, weight_decay=0.0)
        opt.step()

        arrays[f"{name}/loss"] = np.array(terms["loss"].item())
        arrays[f"{name}/rot_mse"] = terms["rot_mse"].detach().numpy()
        if cfg["semantic_guidance"]:
            arrays[f"{name}/text_cosine"] = np.array(terms["text_cosine"].item())
        arrays[f"{name}/grad_digest"] = np.stack([digest(ref_grads[n]) for n, _ in trainable])
        arrays[f"{name}/param_after_digest"] = np.stack([digest(p) for _, p in trainable])
        for n in ("seqTransEncoder.layers.0.norm1.weight", "seqTransEncoder.layers.7.linear2.bias",
                  "seqTransEncoder.layers.3.self_attn.in_proj_bias"):
            arrays[f"{name}/grad/{n}"] = ref_grads[n].numpy()
        arrays[f"{name}/grad_slice/layers.0.linear1.weight"] = ref_grads["seqTransEncoder.layers.0.linear1.weight"][:8, :64].numpy()
        arrays[f"{name}/grad_slice/layers.7.self_attn.in_proj_weight"] = \
            ref_grads["seqTransEncoder.layers.7.self_attn.in_proj_weight"][510:518, :64].numpy()
        arrays["param_names"] = np.array([n for n, _ in trainable])

    # ---- MotionEncoder.forward alone (mu + d mu / d x), with ragged lengths
    from model.mdm_forstyledataset import MotionEncoder as RefMenc  # noqa: F401
    model, _, _ = ref_mu.creat_ddpm_ddim_diffusion(args, ModelClass=StyleDiffusion, timestep_respacing="ddim20")
    model.eval()
    x = inp["x_start"].clone().requires_grad_(True)
    y = {"mask": inp["frame_mask_t2m"][:, None, None, :], "text": inp["texts_t2m"]}
    mu, feat = model.motion_enc(x, y)
    gw = torch.Generator().manual_seed(3)
    wmu = torch.randn(mu.shape, generator=gw)
    (mu * wmu).sum().backward()
    omu = OF.motion_encoder_forward(front, menc, inp["x_start"], inp["frame_mask_t2m"])
    print(f"[menc] oracle vs reference mu: {relerr(omu, mu.detach()):.2e}")
    assert relerr(omu, mu.detach()) < 2e-5
    arrays["menc/mu"] = mu.detach().numpy()
    arrays["menc/w"] = wmu.numpy()
    arrays["menc/dx"] = x.grad.numpy()

    np.savez_compressed(os.path.join(OUT, "finetune.npz"), **arrays)
    print("wrote", os.path.join(OUT, "finetune.npz"), os.path.getsize(os.path.join(OUT, "finetune.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
