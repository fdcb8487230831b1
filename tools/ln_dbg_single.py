#!/usr/bin/env python
"""One eager launch of the LN GEMM with the debug-stamp buffer set (memcheck target)."""
import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mst_b200 import _lib as L  # noqa: E402
m, n, k = 25216, 512, int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = "cuda:0"
lib = L.load()
g = torch.Generator().manual_seed(1)
a = torch.randn(m, k, generator=g).to(dev).bfloat16()
w = (torch.randn(n, k, generator=g) / math.sqrt(k)).to(dev).bfloat16()
bias = torch.randn(n, generator=g).to(dev)
res = torch.randn(m, n, generator=g).to(dev).bfloat16()
out = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
s = torch.cuda.current_stream().cuda_stream
run = lambda: L.check(lib.mst_test_gemm_epi_bf16(2, a.data_ptr(), w.data_ptr(), bias.data_ptr(), res.data_ptr(), bias.data_ptr(), bias.data_ptr(), out.data_ptr(), m, n, k, s))
run(); torch.cuda.synchronize(); print("plain ok")
dbg = torch.zeros(16 * 1024, dtype=torch.int64, device=dev)
torch.cuda.synchronize()
lib.mst_test_set_gemm_debug(dbg.data_ptr())
run(); torch.cuda.synchronize(); print("dbg ok")
lib.mst_test_set_gemm_debug(None)
print(torch.nonzero(dbg.cpu()).flatten()[-5:])
