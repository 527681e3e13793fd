#!/usr/bin/env python
"""tools/ncu_opmix.py <source-page.csv> <DOF> -- dynamic opcode mix of the first kernel in an ncu
`--page source --csv` export: warp instructions executed per opcode, per DOF (x32 = thread instructions)."""
import collections
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
dof = float(sys.argv[2])
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
isrc, iex = hdr.index("Source"), hdr.index("Instructions Executed")
mix = collections.Counter()
for r in rows[hi + 1:]:
    if not r or not r[0].startswith("0x"):
        if mix:
            break
        continue
    toks = r[isrc].split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    mix[op.split(".")[0]] += int(r[iex])
tot = sum(mix.values())
print(f"total warp instructions {tot}  = {tot * 32 / dof:.1f} thread instructions per DOF")
for op, n in mix.most_common(40):
    print(f"  {op:12s} {n:12d}  {n * 32 / dof:7.2f} /DOF  {100 * n / tot:5.1f}%")
