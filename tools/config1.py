#!/usr/bin/env python
"""BASELINE configs[0]: ONE stylexia_posrot inpainting trajectory at B=1 -
  (a) the literal demo path (sample/demo_style_transfer.py:244-258): ddim20, skip_timesteps=14 -> 6 DDIM steps, T=76,
      init_image = content motion, root_horizontal inpainting of the style motion, no CFG wrapper;
  (b) the full 1000-step DDPM p_sample_loop at B=1, T=196 with CFG + inpainting
on the GPU path (both submission modes: whole-trajectory graph / one graph replay per step) and on the CPU oracle port.
    python tools/config1.py [--no-cpu] > gpurun_out/config1.json"""
import argparse
import json
import os
import sys
import time

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402


def gpu_time(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
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
    out = {"config": "BASELINE configs[0]: B=1 stylexia_posrot inpainting trajectory", "gpu": {}, "cpu_port": {}}
    g = torch.Generator().manual_seed(1)
    for mode in ("full", "step"):
        # (a) demo path
        T = 76
        shape = (1, 181, 1, T)
        content, style = torch.randn(shape, generator=g).to(dev), torch.randn(shape, generator=g).to(dev)
        kw = {"y": {"text": ["a person walks"], "text_feat": torch.randn(1, 512, generator=g).to(dev),
                    "mask": torch.ones(1, 1, 1, T, device=dev), "lengths": torch.tensor([T], device=dev),
                    "scale": torch.ones(1, device=dev) * 2.5, "inpainted_motion": style,
                    "inpainting_mask": torch.from_numpy(get_inpainting_mask("root_horizontal", shape)).float().to(dev)}}
        d = mu.create_gaussian_diffusion(bench.Args(), mu.InpaintingGaussianDiffusion, timestep_respacing="ddim20")
        d.rng, d.philox_seed, d.trajectory_graph = "philox", 3, mode
        demo = lambda: d.ddim_sample_loop(model, shape, clip_denoised=False, model_kwargs=kw, skip_timesteps=14,
                                          init_image=content, progress=False)
        ms_demo = gpu_time(demo, 20)
        # (b) 1000-step CFG trajectory at T=196
        T = 196
        shape = (1, 181, 1, T)
        kw2 = {"y": {"text": ["a person walks"], "text_feat": torch.randn(1, 512, generator=g).to(dev),
                     "mask": torch.ones(1, 1, 1, T, device=dev), "lengths": torch.tensor([T], device=dev),
                     "scale": torch.ones(1, device=dev) * 2.5, "inpainted_motion": torch.randn(shape, generator=g).to(dev),
                     "inpainting_mask": torch.from_numpy(get_inpainting_mask("root_horizontal", shape)).float().to(dev)}}
        d2 = mu.create_gaussian_diffusion(bench.Args(), mu.InpaintingGaussianDiffusion)
        d2.rng, d2.philox_seed, d2.trajectory_graph = "philox", 3, mode
        cfg = ClassifierFreeSampleModel(model)
        full = lambda: d2.p_sample_loop(cfg, shape, clip_denoised=False, model_kwargs=kw2)
        ms_full = gpu_time(full, 3)
        out["gpu"][mode] = {"demo_6_ddim_steps_T76_ms": ms_demo, "demo_ms_per_step": ms_demo / 6,
                            "ddpm1000_cfg_T196_ms": ms_full, "ddpm1000_ms_per_step": ms_full / 1000,
                            "ddpm1000_frames_per_s": 196 / (ms_full * 1e-3)}
    if not a.no_cpu:
        sec, threads = bench.cpu_port_step_seconds(1, 196, 3)
        out["cpu_port"] = {"cores": threads, "ddpm1000_cfg_T196_s_per_step": sec, "ddpm1000_frames_per_s": 196 / (1000 * sec),
                           "sample": "3 CFG + inpainting denoise steps at B=1, T=196, extrapolated x1000"}
        # demo path on the CPU port: 6 steps without CFG at T=76
        from oracle import denoiser as OD
        from oracle.weights import mdm_state_dict
        state = mdm_state_dict(181, seed=0)
        x = torch.randn(1, 181, 1, 76)
        feat = torch.randn(1, 512)
        with torch.no_grad():
            OD.mdm_forward(state, x, torch.tensor([300]), feat)
            t0 = time.perf_counter()
            for k in range(6):
                OD.mdm_forward(state, x, torch.tensor([300 - 50 * k]), feat)
            out["cpu_port"]["demo_6_steps_T76_s"] = time.perf_counter() - t0
    os.write(real_stdout, (json.dumps(out) + "\n").encode())


if __name__ == "__main__":
    main()
