#!/usr/bin/env python
"""Graph-replayed timing of the attention kernel: python tools/attn_time.py [n_seqs S]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mst_b200 import _lib as L  # noqa: E402
from mst_b200 import engine as K  # noqa: E402

n_seqs, S = (int(v) for v in sys.argv[1:3]) if len(sys.argv) >= 3 else (128, 197)
dev = "cuda:0"
lib = L.load()
eng = K.Engine(n_feats=181, precision="bf16", device=dev)
qkv = torch.randn(n_seqs * S, 1536, device=dev).bfloat16()
out = torch.empty(n_seqs * S, 512, device=dev, dtype=torch.bfloat16)
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    run = lambda: L.check(lib.mst_test_attention_bf16(eng._h, qkv.data_ptr(), out.data_ptr(), n_seqs, S, None, 0, side.cuda_stream))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, stream=side):
        for _ in range(20):
            run()
gr.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
gr.replay()
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 20 * 1e3
flops = n_seqs * 4 * 4.0 * S * S * 128
print(f"attention n_seqs={n_seqs} S={S}: {us:.1f} us per launch = {flops / us / 1e6:.0f} TFLOP/s")

dbg = torch.zeros(8 * 1024, dtype=torch.int64, device=dev)
lib.mst_test_set_gemm_debug(dbg.data_ptr())
L.check(lib.mst_test_attention_bf16(eng._h, qkv.data_ptr(), out.data_ptr(), n_seqs, S, None, 0, torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
lib.mst_test_set_gemm_debug(None)
d = dbg.cpu().view(8, 1024)
mma = [int(v) for v in d[1] if v != 0]
t0 = mma[0]
print("MMA thread stamps (QK issue / PV issue, in issue order), cycles since first:", [v - t0 for v in mma[:24]])
for grp in range(2):
    e = [int(v) - t0 for v in d[2 + grp] if v != 0]
    print(f"softmax group {grp} (warp quad 0 lane 0): per item: top | s_full wait | pass1 (max) | pass2 (exp, P) | o_full wait | O drain+store")
    nst = int(os.environ.get("MST_ATTN_STAMPS", "7"))
    for i in range(0, len(e) - nst + 1, nst):
        q = e[i:i + nst]
        print(f"   item {i // 7}: top {q[0]:7d} | " + " | ".join(f"{b - a:5d}" for a, b in zip(q, q[1:])) + f" | total {q[-1] - q[0]:6d}")
