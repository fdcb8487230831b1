#!/usr/bin/env python
"""Top stall sites per kernel from `ncu -i X.ncu-rep --page source --csv` output (SASS view).
usage: ncu_src_top.py file.csv [top_n] [kernel_index]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
only = int(sys.argv[3]) if len(sys.argv) > 3 else None
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
for k in range(len(starts) - 1):
    if only is not None and k != only:
        continue
    blk = rows[starts[k]:starts[k + 1]]
    hdr = blk[1]
    data = [r for r in blk[2:] if len(r) == len(hdr)]
    isrc, ismp, iex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
    stall = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    tot = sum(int(r[ismp]) for r in data) or 1
    print(f"== kernel {k}: {blk[0][1][:90]} samples={tot} sass_lines={len(data)}")
    agg = {}
    for r in data:
        for i in stall:
            agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i])
    print("  ", sorted(agg.items(), key=lambda kv: -kv[1])[:8])
    for idx, r in sorted(enumerate(data), key=lambda ir: -int(ir[1][ismp]))[:n]:
        st = sorted([(int(r[i]), hdr[i][6:]) for i in stall if int(r[i]) > 0], reverse=True)[:3]
        print(f"{idx:5d} {int(r[ismp]):6d} {100 * int(r[ismp]) / tot:5.1f}% ex={r[iex]:>8s} {r[isrc].strip()[:64]:64s} {st}")
