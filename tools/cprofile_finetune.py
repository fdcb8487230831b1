#!/usr/bin/env python
"""cProfile of the host side of the finetune step (diagnostic)."""
import cProfile
import os
import pstats
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tools"))
from bench_finetune import build  # noqa: E402

sg = int(sys.argv[1]) if len(sys.argv) > 1 else 0
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
np.random.seed(0)
loop, batch = build(dev, 64, 76, sg, "bf16")
for _ in range(4):
    loop.run_step(*batch)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    loop.run_step(*batch)
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(45)
st.sort_stats("cumulative").print_stats(45)
