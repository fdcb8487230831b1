#!/usr/bin/env python
"""Time the native CLIP text tower (scope row N1): B captions x 77 tokens, CUDA events after warm-up.
Algorithmic flops per caption: 12 layers x (8 S w^2 + 4 S^2 w + 4 S w ff) + 2 w d_out, S=77, w=512, ff=2048."""
import argparse
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
import conftest  # noqa: F401,E402  (registers the mst_b200 alias)
from mst_b200.model.clip_text import CLIPTextTower  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    S, w, ff, L = 77, 512, 2048, 12
    flops = a.batch * (L * (8 * S * w * w + 4 * S * S * w + 4 * S * w * ff) + 2 * w * 512)
    tok = torch.zeros(a.batch, 77, dtype=torch.int32, device="cuda")
    tok[:, 0], tok[:, 1:9], tok[:, 9] = 49406, 1234, 49407
    for prec in ("fp32", "bf16"):
        tower = CLIPTextTower(precision=prec).cuda()
        for _ in range(3):
            tower.encode_text(tok)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(a.iters):
            tower.encode_text(tok)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.iters
        print(json.dumps({"metric": "clip_text_encode_ms", "precision": prec, "batch": a.batch, "value": round(ms, 4),
                          "tflops": round(flops / ms / 1e9, 2), "captions_per_s": round(a.batch / ms * 1e3)}))


if __name__ == "__main__":
    main()
