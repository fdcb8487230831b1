#!/usr/bin/env python
"""Opcode histogram per kernel of the built library (cuobjdump -sass): the Blackwell-native instructions
(UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA, UTCBAR = tcgen05.commit,
SYNCS = mbarrier) next to the legacy ones (HMMA = mma.sync / wmma): zero in every sampler kernel; the only HMMA in the
library are the tf32 m16n8k8 products of the short-sequence TRAINING attention (attn_small_*<..., true>: 80 x 80 x 128
products held in shared memory by one CTA per (sequence, head); DESIGN.md section 4, round-2 finetune table).
    python tools/sass_histogram.py > profiles/rNN_sass_opcode_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(REPO, "diffusion-based-motion-style-transfer_b200", "libmst_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
WATCH = ["UTCHMMA", "UTCHMMA.2CTA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "SYNCS", "MUFU.EX2",
         "FFMA2", "FADD2", "FMUL2", "F2FP", "HMMA", "HGMMA", "LDGSTS", "REDUX", "STG.E.128", "LDG.E.128"]
cur, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur)
        hist[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        hist[cur]["_total"] += 1
        for w in WATCH:
            if op == w or op.startswith(w + ".") or (w.count(".") and op.startswith(w)):
                hist[cur][w] += 1
print(f"{os.path.basename(lib)}: SASS opcode counts per kernel (static instruction counts)")
print(f"{'kernel':58s} {'instr':>6s} " + " ".join(f"{w:>9s}" for w in WATCH))
tot = collections.Counter()
for k, c in hist.items():
    if not (c["UTCHMMA"] or c["LDTM"] or c["UTMALDG"] or c["HMMA"] or "tc_" in k or "update" in k):
        continue
    print(f"{k[:58]:58s} {c['_total']:6d} " + " ".join(f"{c[w]:9d}" for w in WATCH))
    tot.update(c)
print(f"{'ALL listed kernels':58s} {tot['_total']:6d} " + " ".join(f"{tot[w]:9d}" for w in WATCH))
print(f"legacy tensor-core opcodes in the whole library: HMMA {sum(c['HMMA'] for c in hist.values())}, "
      f"HGMMA {sum(c['HGMMA'] for c in hist.values())}")
print("kernels with HMMA (mma.sync): " + (", ".join(k for k, c in hist.items() if c["HMMA"]) or "none") +
      "  [training attention of the finetune step only; no sampler kernel]")
