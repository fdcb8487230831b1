#!/usr/bin/env python
"""clock64 timeline of cluster 0 of the pair GEMM (QKV shape by default): where do the producer, the MMA
issuer and the epilogue spend their cycles?   python tools/gemm_timeline.py [epi m n k]"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mst_b200 import _lib as L  # noqa: E402

epi, m, n, k = (int(v) for v in sys.argv[1:5]) if len(sys.argv) >= 5 else (0, 25216, 1536, 512)
dev = "cuda:0"
lib = L.load()
g = torch.Generator().manual_seed(1)
a = torch.randn(m, k, generator=g).to(dev).bfloat16()
w = (torch.randn(n, k, generator=g) / math.sqrt(k)).to(dev).bfloat16()
bias = torch.randn(n, generator=g).to(dev)
res = torch.randn(m, n, generator=g).to(dev).bfloat16()
out = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
s = torch.cuda.current_stream().cuda_stream


def run():
    L.check(lib.mst_test_gemm_epi_bf16(epi, a.data_ptr(), w.data_ptr(), bias.data_ptr(), res.data_ptr(), bias.data_ptr(),
                                       bias.data_ptr(), out.data_ptr(), m, n, k, s))


SINGLE = os.environ.get('MST_TL_SINGLE') == '1'
for _ in range(3):
    run()
torch.cuda.synchronize()
# the stamped launch runs first (a stamped eager launch AFTER the graph sections below faulted; cause unknown)
dbg = torch.zeros(16 * 1024, dtype=torch.int64, device=dev)
lib.mst_test_set_gemm_debug(dbg.data_ptr())
run()
torch.cuda.synchronize()
lib.mst_test_set_gemm_debug(None)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    run()
e1.record()
torch.cuda.synchronize()
eager_us = e0.elapsed_time(e1) / 10 * 1e3
# the same launches replayed from a CUDA graph: no host launch cost between kernels
gr = torch.cuda.CUDAGraph()
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    s = side.cuda_stream
    run()
    torch.cuda.synchronize()
    with torch.cuda.graph(gr, stream=side):
        for _ in range(20):
            run()
s = torch.cuda.current_stream().cuda_stream
gr.replay()
torch.cuda.synchronize()
e0.record()
gr.replay()
e1.record()
torch.cuda.synchronize()
flops = 2.0 * m * n * k
graph_us = e0.elapsed_time(e1) / 20 * 1e3
print(f"epi={epi} m={m} n={n} k={k}: eager {eager_us:.1f} us, graph {graph_us:.1f} us per launch = {flops / graph_us / 1e6:.0f} TFLOP/s")
if epi == 2:
    # gap between two consecutive launches inside one graph: wall-clock exit of launch 1 vs entry of launch 2
    bufs = [torch.zeros(16 * 1024, dtype=torch.int64, device=dev) for _ in range(3)]
    g2 = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        s = side.cuda_stream
        with torch.cuda.graph(g2, stream=side):
            for b in bufs:
                lib.mst_test_set_gemm_debug(b.data_ptr())
                run()
        lib.mst_test_set_gemm_debug(None)
    s = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        g2.replay()
    torch.cuda.synchronize()
    ws = []
    for b in bufs:
        w = b.cpu()[6 * 1024:6 * 1024 + 3 * 148].view(148, 3)
        ws.append(w[w[:, 0] > 0])
    t0g = int(ws[0][:, 0].min())
    for i, w in enumerate(ws):
        print(f"  launch {i}: first entry {int(w[:,0].min()) - t0g} ns, last entry {int(w[:,0].max()) - t0g}, first exit {int(w[:,1].min()) - t0g}, last exit {int(w[:,1].max()) - t0g}")
_nz = torch.nonzero(dbg.cpu()[7 * 1024:]).flatten()
if len(_nz):
    print("WARNING: debug stamps beyond 7*1024:", (_nz[:8] + 7 * 1024).tolist(), "count", len(_nz))
if epi == 2:
    w = dbg.cpu()[6 * 1024:6 * 1024 + 3 * 148].view(148, 3)
    w = w[w[:, 0] > 0]
    t_first = int(w[:, 0].min())
    print(f"CTA wall-clock (ns since first CTA entry): entries min/median/max = {int(w[:,0].min()) - t_first}/{int(w[:,0].median()) - t_first}/{int(w[:,0].max()) - t_first}; "
          f"exits min/median/max = {int(w[:,1].min()) - t_first}/{int(w[:,1].median()) - t_first}/{int(w[:,1].max()) - t_first}")
    dur = (w[:, 1] - w[:, 0]).float()
    print(f"  per-CTA duration ns: min {dur.min():.0f} median {dur.median():.0f} max {dur.max():.0f}; CTA0 {int(w[0,1]-w[0,0])}")
d = dbg.cpu()[:6 * 1024 - 1024 * 0][:3 * 2 * 1024].view(3, 2, 1024)
kb = k // 64
t0 = int(d[1, 0, 0])
rel = lambda v: int(v) - t0
prod = [[rel(v) for v in d[0, r] if v != 0] for r in range(2)]
mma = [rel(v) for v in d[1, 0] if v != 0]
epi_t = [[rel(v) for v in d[2, r] if v != 0] for r in range(2)]
per = 2 + kb
print("MMA issuer (leader): per tile [start, after tempty wait, after full wait of each k-block]")
for i in range(0, len(mma), per):
    row = mma[i:i + per]
    print(f"  tile {i // per}: start {row[0]:7d} tempty+{row[1] - row[0]:5d} kblocks " + " ".join(f"{b - a_:4d}" for a_, b in zip(row[1:], row[2:])),
          f"| tile total {row[-1] - row[0]:6d}")
if epi == 2:
    print("LN epilogue (warp 2 lane 0) per tile: top | residual wait + read | tfull wait | pass1 | stats exchange | pass2+fence | store+drain+next residual issue")
    for r in range(2):
        e = epi_t[r]
        for i in range(0, len(e), 7):
            q = e[i:i + 7]
            if len(q) == 7:
                print(f"  cta{r} tile {i // 7}: top {q[0]:7d} | " + " | ".join(f"{b - a_:5d}" for a_, b in zip(q, q[1:])) + f" | total {q[6] - q[1]:5d}")
    sys.exit(0)
print("epilogue (warp 2 lane 0) per tile: top | tfull wait | first 3 column steps (tmem released) | last step + stores | total after tfull")
for r in range(2):
    e = epi_t[r]
    for i in range(0, len(e), 4):
        q = e[i:i + 4]
        if len(q) == 4:
            print(f"  cta{r} tile {i // 4}: top {q[0]:7d} | " + " | ".join(f"{b - a_:5d}" for a_, b in zip(q, q[1:])) + f" | total {q[3] - q[1]:5d}")
print("producer: time after each empty wait (per k-block), deltas")
for r in range(2):
    pr = prod[r]
    print(f"  cta{r}: first {pr[0]}", " ".join(str(b - a_) for a_, b in zip(pr[:40], pr[1:41])))
