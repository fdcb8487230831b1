#!/usr/bin/env python
"""A few eager launches of the attention kernel (ncu target): python tools/attn_eager.py [n_seqs S n]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mst_b200 import _lib as L  # noqa: E402
from mst_b200 import engine as K  # noqa: E402
n_seqs, S, n = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (128, 197, 4)
dev = "cuda:0"
lib = L.load()
eng = K.Engine(n_feats=181, precision="bf16", device=dev)
qkv = torch.randn(n_seqs * S, 1536, device=dev).bfloat16()
out = torch.empty(n_seqs * S, 512, device=dev, dtype=torch.bfloat16)
for _ in range(n):
    L.check(lib.mst_test_attention_bf16(eng._h, qkv.data_ptr(), out.data_ptr(), n_seqs, S, None, 0, torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
