#!/usr/bin/env python
"""2+ ranks (torchrun): the overlapped gradient exchange of the finetune step (t2m gradient all-reduced asynchronously
while the style term runs) must give the same parameters as the blocking all-reduce of the whole arena.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/check_overlap_equiv.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tools"))
from bench_finetune import build  # noqa: E402

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
os.dup2(2, 1)
dist.init_process_group("nccl", device_id=dev)
res = {}
for mode in ("1", "0"):
    os.environ["MST_OVERLAP_ALLREDUCE"] = mode
    np.random.seed(0)
    torch.manual_seed(0)
    torch.cuda.manual_seed(0)
    loop, batch = build(dev, 64, 76, 1, "fp32")
    loop.model.mst_train_dropout = 0.0  # deterministic: the two modes must see the same forward
    for _ in range(3):
        loop.run_step(*batch)
    torch.cuda.synchronize()
    res[mode] = (loop.mp_trainer.flat.train_params.clone(), float(loop.last_losses["loss"]), loop.last_overlap())
a, b = res["1"][0], res["0"][0]
diff = float((a - b).abs().max())
moved = float((a - b).abs().max() / 1e-4)
if rank == 0:
    print(f"overlapped vs blocking after 3 steps: max |param diff| = {diff:.3e} ({moved:.3e} of one lr step); "
          f"loss {res['1'][1]:.6f} vs {res['0'][1]:.6f}; overlap window / exposed wait (ms) = {res['1'][2]}", file=sys.stderr)
    assert diff < 2e-6, diff
dist.destroy_process_group()
