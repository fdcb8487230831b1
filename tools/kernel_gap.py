#!/usr/bin/env python
"""Dead time between consecutive kernels of one CUDA graph, from %globaltimer stamps inside the kernels:
   python tools/kernel_gap.py <epi: 0 pair-bias | 1 pair-gelu | 2 ln-cluster | 9 single-cta fp32> m n k"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mst_b200 import _lib as L  # noqa: E402

epi, m, n, k = (int(v) for v in sys.argv[1:5])
dev = "cuda:0"
lib = L.load()
g = torch.Generator().manual_seed(1)
a = torch.randn(m, k, generator=g).to(dev).bfloat16()
w = (torch.randn(n, k, generator=g) / math.sqrt(k)).to(dev).bfloat16()
bias = torch.randn(n, generator=g).to(dev)
res = torch.randn(m, n, generator=g).to(dev).bfloat16()
out = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
out32 = torch.empty(m, n, device=dev)
side = torch.cuda.Stream()


def run():
    s = side.cuda_stream
    if epi == 9:
        L.check(lib.mst_test_gemm_bf16(a.data_ptr(), w.data_ptr(), bias.data_ptr(), out32.data_ptr(), m, n, k, s))
    else:
        L.check(lib.mst_test_gemm_epi_bf16(epi, a.data_ptr(), w.data_ptr(), bias.data_ptr(), res.data_ptr(), bias.data_ptr(),
                                           bias.data_ptr(), out.data_ptr(), m, n, k, s))


bufs = [torch.zeros(9 * 1024, dtype=torch.int64, device=dev) for _ in range(4)]
with torch.cuda.stream(side):
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, stream=side):
        for b in bufs:
            lib.mst_test_set_gemm_debug(b.data_ptr())
            run()
    lib.mst_test_set_gemm_debug(None)
for _ in range(3):
    gr.replay()
torch.cuda.synchronize()
ws = []
for b in bufs:
    w_ = b.cpu()[6 * 1024:6 * 1024 + 3 * 148].view(148, 3)
    ws.append(w_[w_[:, 0] > 0])
t0 = int(ws[0][:, 0].min())
print(f"epi={epi} m={m} n={n} k={k} PDL={os.environ.get('MST_PDL', '1')}: per launch [first entry, last entry | first work-end, "
      f"last work-end | last exit] ns")
prev_exit = None
for i, w_ in enumerate(ws):
    fe, le = int(w_[:, 0].min()) - t0, int(w_[:, 0].max()) - t0
    fw, lw, lx = int(w_[:, 1].min()) - t0, int(w_[:, 1].max()) - t0, int(w_[:, 2].max()) - t0
    gap = "" if prev_exit is None else f"  gap after previous kernel's last exit: {fe - prev_exit} ns"
    print(f"  launch {i}: {fe:7d} {le:7d} | {fw:7d} {lw:7d} | {lx:7d}{gap}")
    prev_exit = lx
if epi == 9:
    d = bufs[1].cpu()[7 * 1024:7 * 1024 + 2 * 148].view(148, 2)
    d = d[d[:, 0] > 0]
    dd = (d[:, 1] - d[:, 0]).float()
    print(f"  tcgen05.dealloc duration per CTA (ns): min {dd.min():.0f} median {dd.median():.0f} max {dd.max():.0f}; "
          f"pre-dealloc stamps rel t0: min {int(d[:,0].min()) - t0} max {int(d[:,0].max()) - t0}")
