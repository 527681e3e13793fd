#!/usr/bin/env python
"""tools/show_bench.py <bench json line file> -- human-readable digest of one bench.py line."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = d.get("roofline") or {}
print(f"{d['config']['workload']}  n_gpus {d['n_gpus']}  value {d['value']:.1f} {d['unit']}  {d['ms_per_step'] * 1e3:.2f} us/step  "
      f"hbm frac {r.get('frac', 0):.3f}  fp32 frac {(r.get('secondary') or {}).get('frac', 0):.3f}  e2e {(d.get('e2e') or {}).get('value', 0):.2f}")
for k, v in (d.get("points") or {}).items():
    if "error" in v:
        print(f"  {k:26s} ERROR {v['error']}")
    else:
        sec = v.get("secondary", {}).get("frac")
        print(f"  {k:34s} {v['value']:7.1f} GDOF/s  {v['ms_per_step'] * 1e3:8.2f} us  hbm {v['roofline']['frac']:.3f}"
              + (f"  fp32 {sec:.3f}" if sec is not None else "") + f"  {v.get('launch', '')}")
if d.get("slab"):
    s = d["slab"]
    print("  slab:", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in s.items() if k not in ("desc", "l2", "roofline_per_gpu")})
print("  cpu_baseline:", d.get("cpu_baseline"))
print("  train:", d.get("train"))
print("  clocks:", d.get("clocks"))
