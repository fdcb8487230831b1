#!/usr/bin/env python
"""Where does the host time of one finetune step go?  cProfile of loop.run_step (top cumulative / own time) next to the
kernel-time table of tools/bench_finetune.py --profile.   python tools/finetune_hostprof.py [--sg 1] [--batch 64]"""
import argparse
import cProfile
import io
import os
import pstats
import sys
import time

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tools"))
import bench_finetune as bf  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sg", type=int, default=1)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--frames", type=int, default=76)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--dump", default="", help="write the (kernel, ms) records of one step here and stop (use MST_TRAIN_GRAPH=0)")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    from mst_b200 import engine as K
    loop, batch = bf.build(dev, a.batch, a.frames, a.sg)
    for _ in range(5):
        loop.run_step(*batch)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        loop.run_step(*batch)
    issue = (time.perf_counter() - t0) / a.steps * 1e3
    torch.cuda.synchronize()
    total = (time.perf_counter() - t0) / a.steps * 1e3
    print(f"sg={a.sg} B={a.batch}: host issue {issue:.3f} ms/step, wall {total:.3f} ms/step")
    with K.profile(cap=8192) as p:
        loop.run_step(*batch)
    agg = {}
    for name, t in p.records:
        c = agg.setdefault(name, [0, 0.0])
        c[0] += 1
        c[1] += t
    tot = sum(v[1] for v in agg.values())
    print(f"kernel time of one step (serialised, events around each launch): {tot:.3f} ms in {sum(v[0] for v in agg.values())} launches")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
        print(f"   {k:40s} x{v[0]:3d}  {v[1]:.4f} ms")
    if a.dump:
        import json
        with open(a.dump, "w") as f:
            json.dump([[n, t] for n, t in p.records], f)
        return
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(a.steps):
        loop.run_step(*batch)
    pr.disable()
    torch.cuda.synchronize()
    for key in ("cumulative", "tottime"):
        s = io.StringIO()
        pstats.Stats(pr, stream=s).sort_stats(key).print_stats(45)
        print(s.getvalue()[:9000])


if __name__ == "__main__":
    main()
