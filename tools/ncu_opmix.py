"""Summarise an ncu source page (ncu -i X.ncu-rep --page source --csv > src.csv): executed warp instructions by
opcode, and the hottest instruction ranges.   python tools/ncu_opmix.py src.csv [row_ops]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
body = rows[2:]
norm = float(sys.argv[2]) if len(sys.argv) > 2 else None
ops = collections.Counter()
samples = collections.Counter()
tot = 0
for r in body:
    if len(r) < len(hdr):
        continue
    src = r[ci["Source"]].strip()
    n = int(r[ci["Instructions Executed"]] or 0)
    s = int(r[ci["# Samples"]] or 0)
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
    op = op.split(".")[0]
    ops[op] += n
    samples[op] += s
    tot += n
print(f"total warp instructions: {tot:.4g}" + (f"  = {tot / norm:.1f} per row-op" if norm else ""))
for op, n in ops.most_common(28):
    print(f"  {op:12s} {n:14d} {100 * n / tot:5.1f}%  samples {samples[op]:7d}" + (f"  {n / norm:6.2f}/row-op" if norm else ""))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.Counter()
for r in body:
    if len(r) < len(hdr):
        continue
    for h in stalls:
        agg[h] += int(r[ci[h]] or 0)
st = sum(agg.values())
print("stall samples:", ", ".join(f"{h[6:]}={100 * v / st:.1f}%" for h, v in agg.most_common(8)))
