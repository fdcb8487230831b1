#!/usr/bin/env python
"""Golden for the HEADLINE shape (BASELINE configs[1]: B=64, F=181, T=196, CFG scale 2.5 + root_horizontal
inpainting): ONE DDPM step (t = 999) of the REAL reference's p_sample_loop, plus the oracle check at that size.

    python tests/golden/make_golden_b64.py        # build container only (needs /root/reference)

The full x0 tensor is 9 MB, so the fixture keeps four samples in full (0, 21, 42, 63) and, for all 64 samples, three
digests of the x0 prediction (sum, L2 norm, dot product with a seeded probe) - enough to catch an error in any sample.
Inputs are regenerated from seeds by the tests (tests/golden/b64_inputs.py)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402  (shims, Args, the reference import path)
from b64_inputs import b64_inputs, b64_digests, KEEP  # noqa: E402
from oracle import denoiser as OD, sampler as OS, schedule as OSch  # noqa: E402
from oracle.weights import NoiseTape, mdm_state_dict  # noqa: E402


def main():
    torch.set_grad_enabled(False)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    MG.install_shims()
    from utils import model_util as ref_mu
    from diffusion.inpainting_gaussian_diffusion import InpaintingGaussianDiffusion
    from model.cfg_sampler import ClassifierFreeSampleModel
    from model.mdm_forstyledataset import MDM
    from data_loaders import stylexia_posrot_utils as ref_masks
    import torch as th

    args = MG.Args()
    state = mdm_state_dict(n_feats=181, seed=0)
    ref_model = MDM(**ref_mu.get_transfer_args(args))
    ref_model.load_state_dict(state, strict=False)
    ref_model.train(False)
    inp = b64_inputs()
    shape = inp["shape"]
    mask = torch.from_numpy(ref_masks.get_inpainting_mask("root_horizontal", shape)).float()
    assert torch.equal(mask, inp["mask"])
    B, T = shape[0], shape[3]
    yk = {"y": {"text": inp["texts"], "mask": torch.ones(B, 1, 1, T), "lengths": torch.tensor([T] * B), "scale": inp["scale"],
                "inpainted_motion": inp["x_inp"], "inpainting_mask": mask}}
    d = ref_mu.create_gaussian_diffusion(args, InpaintingGaussianDiffusion)
    tape = NoiseTape(3)
    orig_randn, orig_like = th.randn, th.randn_like
    th.randn = lambda *s, **k: tape.draw(s[0] if len(s) == 1 and isinstance(s[0], (tuple, list, torch.Size)) else s)
    th.randn_like = lambda a, **k: tape.draw(a.shape)
    try:
        xs = d.p_sample_loop(ClassifierFreeSampleModel(ref_model), shape, clip_denoised=False, model_kwargs=yk,
                             stop_timesteps=999, dump_all_xstart=True)
    finally:
        th.randn, th.randn_like = orig_randn, orig_like
    assert len(xs) == 1
    x0 = xs[0]
    sch = OSch.Schedule(OSch.cosine_betas(1000))
    _, xs_o = OS.sample_loop(sch, lambda xx, tt: OD.cfg_forward(state, xx, tt, inp["feat"], inp["scale"]), shape, NoiseTape(3),
                             mask=mask, x_inp=inp["x_inp"], stop_timesteps=999)
    err = float((xs_o[0] - x0).abs().max() / x0.abs().max())
    print(f"oracle vs reference, B=64 T=196 one CFG + inpainting step: x0 err {err:.2e}")
    assert err < 5e-5
    out = {"x0_keep": x0[KEEP].numpy()}
    out.update({k: v.numpy() for k, v in b64_digests(x0).items()})
    path = os.path.join(HERE, "b64_step.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
