#!/usr/bin/env python
"""A few eager denoise steps of the BASELINE configs[1] workload (B=64, T=196, CFG + inpainting, bf16) for ncu:
    python tools/profile_step.py [--steps 2] [--batch 64] [--frames 196]
Launch order per step: token0, motion_to_tokens, tc_gemm_inproj, 8 x (qkv, attention, res_ln, ffn1, res_ln),
tc_gemm_outproj, update  = 45 kernels.  Prints the per-kernel CUDA-event table when not under a profiler."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--frames", type=int, default=196)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--eager", action="store_true", help="eager launches only (use under ncu)")
    a = ap.parse_args()
    from mst_b200 import engine as K
    from mst_b200.model.mdm_forstyledataset import MDM
    from mst_b200.utils import model_util as mu
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    torch.manual_seed(0)
    model = MDM(load_clip=False, **mu.get_transfer_args(bench.Args()))
    model.mst_precision = a.precision
    model.to(dev).eval()
    roof, stages = bench.roofline_leg(K, model, dev, a.batch, a.frames, bench.peaks(), 0.0, steady=not a.eager)
    for s in stages:
        print(s)
    print(roof)


if __name__ == "__main__":
    main()
