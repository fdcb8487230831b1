#!/usr/bin/env python
"""The fused update kernel alone at an HBM-bound size (B=2048, T=196: 1.45 GB per launch) - for ncu.
    python tools/profile_update.py [--batch 2048] [--reps 6]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=2048)
    ap.add_argument("--reps", type=int, default=6)
    a = ap.parse_args()
    from mst_b200 import _lib as L
    from mst_b200 import engine as K
    from mst_b200.utils import model_util as mu
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    B, F, T = a.batch, 181, 196
    shape = (B, F, 1, T)
    oc, ou, x, xi = (torch.randn(shape, device=dev) for _ in range(4))
    out = torch.empty_like(x)
    scale = torch.full((B,), 2.5, device=dev)
    mask = torch.zeros(F, device=dev)
    mask[:3] = 1
    tabs = mu.create_gaussian_diffusion(bench.Args(), mu.InpaintingGaussianDiffusion).device_tables(dev)
    t = torch.full((B,), 500, device=dev, dtype=torch.long)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(a.reps):
        if i == a.reps - 1:
            e0.record()
        K.update_step(sampler=L.SAMPLER_DDPM, out_cond=oc, out_uncond=ou, cfg_scale=scale, x_t=x, x_prev=out, mask=mask,
                      x_inpaint=xi, mask_noise=True, clip_denoised=False, t_vec=t, coef1=tabs["c1"], coef2=tabs["c2"],
                      sigma=tabs["sigma"], noise_kind=L.NOISE_PHILOX, philox_seed=3)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"update_kernel B={B}: {ms * 1e3:.1f} us, {20 * B * F * T / ms / 1e6:.0f} GB/s algorithmic (20 B/element)")


if __name__ == "__main__":
    main()
