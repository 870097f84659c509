#!/usr/bin/env python
"""Opcode histogram + stall summary from `ncu -i X.ncu-rep --page source --csv` (first kernel instance)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter(); stall = collections.Counter(); samples = collections.Counter()
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = 0
seen_kernel = 0
for r in rows[2:]:
    if len(r) < len(hdr):
        if r and r[0] == "Kernel Name":
            seen_kernel += 1
            if seen_kernel >= 1: break
        continue
    if r[0] == "Address": continue
    src = r[ix["Source"]].strip()
    toks = src.split()
    if not toks: continue
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = op.split(".")[0]
    n = int(r[ix["Instructions Executed"]] or 0)
    ops[op] += n; tot += n
    samples[op] += int(r[ix["# Samples"]] or 0)
    for c in stall_cols:
        stall[c] += int(r[ix[c]] or 0)
print("total warp instructions", tot)
for op, n in ops.most_common(40):
    print(f"{op:12s} {n:12d} {100*n/tot:6.2f}%  samples {samples[op]}")
ts = sum(stall.values())
print("stalls:")
for c, n in stall.most_common(12):
    print(f"  {c:28s} {100*n/max(ts,1):6.2f}%")
