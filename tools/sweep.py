#!/usr/bin/env python
"""BASELINE configs[4] / configs[2]: respaced 50-step p_sample_loop (CFG + inpainting, bf16) over batch and sequence
length - latency (ms / trajectory) against throughput (frames/s) - plus the fused update kernel alone at an
HBM-resident-free size (its roofline leg: achieved GB/s at B >= 2048).

    python tools/sweep.py [--batches 1,8,64,512,2048] [--frames 60,76,120,196] [--out gpurun_out/sweep.json]
"""
import argparse
import json
import os
import sys
import time

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", default="1,8,64,512,2048")
    ap.add_argument("--frames", default="60,76,120,196")
    ap.add_argument("--respacing", default="50")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    from mst_b200 import _lib as L
    from mst_b200 import engine as K
    from mst_b200.data_loaders.stylexia_posrot_utils import get_inpainting_mask
    from mst_b200.model.cfg_sampler import ClassifierFreeSampleModel
    from mst_b200.model.mdm_forstyledataset import MDM
    from mst_b200.utils import model_util as mu
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    torch.manual_seed(0)
    model = MDM(load_clip=False, **mu.get_transfer_args(bench.Args()))
    model.mst_precision = "bf16"
    model.to(dev).eval()
    cfg = ClassifierFreeSampleModel(model)
    rows = []
    pk = bench.peaks()
    for T in [int(v) for v in a.frames.split(",")]:
        for B in [int(v) for v in a.batches.split(",")]:
            d = mu.create_gaussian_diffusion(bench.Args(), mu.InpaintingGaussianDiffusion, timestep_respacing=a.respacing)
            d.rng = "philox"
            shape = (B, 181, 1, T)
            g = torch.Generator().manual_seed(1)
            x_inp = torch.randn(shape, generator=g).to(dev)
            mask = torch.from_numpy(get_inpainting_mask("root_horizontal", shape)).float().to(dev)
            y = {"y": {"text": ["x"] * B, "text_feat": torch.randn(B, 512, generator=g).to(dev), "scale": torch.full((B,), 2.5, device=dev),
                       "mask": torch.ones(B, 1, 1, T, device=dev), "lengths": torch.full((B,), T), "inpainted_motion": x_inp,
                       "inpainting_mask": mask}}
            for _ in range(2):
                d.p_sample_loop(cfg, shape, clip_denoised=False, model_kwargs=y)
            torch.cuda.synchronize()
            reps = 3 if B <= 512 else 2
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                d.p_sample_loop(cfg, shape, clip_denoised=False, model_kwargs=y)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            n = d.num_timesteps
            fl = 2 * B * bench.model_flops_per_seq(T) * n
            row = {"B": B, "T": T, "steps": n, "ms_per_trajectory": round(ms, 3), "ms_per_denoise_step": round(ms / n, 4),
                   "frames_per_s": round(B * T / (ms / 1e3), 1), "model_tflops": round(fl / (ms / 1e3) / 1e12, 1),
                   "frac_sustained_peak": round(fl / (ms / 1e3) / 1e12 / pk["bf16_tflops_sustained"], 3)}
            rows.append(row)
            print(json.dumps(row), flush=True)
            del d, x_inp, mask, y
            torch.cuda.empty_cache()
    # the fused update kernel alone, HBM-bound sizes (20 B / element: out_c, out_u, x_t, x_inp read, x_{t-1} written)
    upd = []
    for B in (64, 512, 2048, 4096):
        T, F = 196, 181
        shape = (B, F, 1, T)
        oc, ou, x, xi = (torch.randn(shape, device=dev) for _ in range(4))
        out = torch.empty_like(x)
        scale = torch.full((B,), 2.5, device=dev)
        mask = torch.zeros(F, device=dev)
        mask[:3] = 1
        d = mu.create_gaussian_diffusion(bench.Args(), mu.InpaintingGaussianDiffusion)
        tabs = d.device_tables(dev)
        t = torch.full((B,), 500, device=dev, dtype=torch.long)

        def run():
            K.update_step(sampler=L.SAMPLER_DDPM, out_cond=oc, out_uncond=ou, cfg_scale=scale, x_t=x, x_prev=out, mask=mask,
                          x_inpaint=xi, mask_noise=True, clip_denoised=False, t_vec=t, coef1=tabs["c1"], coef2=tabs["c2"],
                          sigma=tabs["sigma"], noise_kind=L.NOISE_PHILOX, philox_seed=3)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record()
        for _ in range(reps):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        nbytes = 20 * B * F * T
        row = {"kernel": "update", "B": B, "bytes": nbytes, "ms": round(ms, 4), "GBps": round(nbytes / ms / 1e6, 1),
               "frac_measured_hbm": round(nbytes / ms / 1e6 / pk["hbm_gbs"], 3), "frac_8TBps": round(nbytes / ms / 1e6 / 8000, 3)}
        upd.append(row)
        print(json.dumps(row), flush=True)
        del oc, ou, x, xi, out
        torch.cuda.empty_cache()
    if a.out:
        with open(a.out, "w") as f:
            json.dump({"sweep": rows, "update_kernel": upd, "peaks": pk}, f, indent=1)


if __name__ == "__main__":
    main()
