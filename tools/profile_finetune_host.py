#!/usr/bin/env python
"""Host-side view of the finetune step (torch.profiler / CUPTI): where the Python thread spends its time and how busy the
GPU is.  Diagnostic only - numbers taken under a profiler are never bench values."""
import argparse
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tools"))
from bench_finetune import build  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sg", type=int, default=0)
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    np.random.seed(0)
    loop, batch = build(dev, 64, 76, a.sg, "bf16")
    for _ in range(4):
        loop.run_step(*batch)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(a.steps):
            loop.run_step(*batch)
        torch.cuda.synchronize()
    ev = prof.key_averages()
    kern = sum(e.self_device_time_total for e in ev) / a.steps / 1e3
    print(f"GPU kernel time per step: {kern:.3f} ms")
    print(ev.table(sort_by="self_cpu_time_total", row_limit=30, max_name_column_width=60))
    print(ev.table(sort_by="self_device_time_total", row_limit=25, max_name_column_width=60))




def synced_breakdown(sg=0, steps=5):
    """GPU-inclusive time of every Engine.forward_train / backward call (sync before and after) vs the whole step."""
    import time
    from mst_b200 import engine as K
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    np.random.seed(0)
    loop, batch = build(dev, 64, 76, sg, "bf16")
    for _ in range(4):
        loop.run_step(*batch)
    torch.cuda.synchronize()
    acc = {}

    def wrap(name):
        fn = getattr(K.Engine, name)

        def timed(self, *a, **k):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = fn(self, *a, **k)
            torch.cuda.synchronize()
            c = acc.setdefault(name, [0, 0.0])
            c[0] += 1
            c[1] += (time.perf_counter() - t0) * 1e3
            return r
        setattr(K.Engine, name, timed)

    for n in ("forward_train", "backward", "motion_encoder_forward", "motion_encoder_backward"):
        wrap(n)
    t0 = time.perf_counter()
    for _ in range(steps):
        loop.run_step(*batch)
    torch.cuda.synchronize()
    tot = (time.perf_counter() - t0) * 1e3 / steps
    print(f"sg={sg}: synced step {tot:.3f} ms; per step:",
          {k: (v[0] // steps, round(v[1] / steps, 3)) for k, v in acc.items()})


if __name__ == "__main__":
    if os.environ.get("MST_SYNCED_BREAKDOWN"):
        synced_breakdown(int(os.environ["MST_SYNCED_BREAKDOWN"]) - 1)
    else:
        main()
