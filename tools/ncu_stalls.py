#!/usr/bin/env python
"""tools/ncu_stalls.py <source-page.csv> [N] -- stall-reason totals and the N hottest SASS lines."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
data = []
for r in rows[hi + 1:]:
    if not r or not r[0].startswith("0x"):
        if data:
            break            # first kernel section only
        continue
    data.append(r)
ia, isrc, isamp, iex = (hdr.index(k) for k in ("Address", "Source", "# Samples", "Instructions Executed"))
names = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
idx = {n: hdr.index(n) for n in names}
tot = sum(int(r[isamp]) for r in data)
print("total samples", tot, " instructions executed", sum(int(r[iex]) for r in data))
for n in names:
    s = sum(int(r[idx[n]]) for r in data)
    if s:
        print(f"{n:28s}{s:8d} {100 * s / tot:5.1f}%")
print()
base = int(data[0][ia], 16)
top = sorted(data, key=lambda r: -int(r[isamp]))[:N]
for r in sorted(top, key=lambda r: int(r[ia], 16)):
    st = {n: int(r[idx[n]]) for n in names if int(r[idx[n]]) > 0}
    main = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(f"{int(r[ia], 16) - base:5x} {r[isamp]:>5s} {r[iex]:>7s} {r[isrc].strip()[:64]:64s} {main}")
