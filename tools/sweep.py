#!/usr/bin/env python
"""tools/sweep.py [workload ...] -- time the fused 3-D launch under a list of env-knob settings in ONE
process (the planners read the environment per call), and cross-check every setting against the
first one (loss rel. difference, gradient max-norm difference).

    python tools/sweep.py --cfg "DN_T3_THREADS=256" --cfg "DN_T3_THREADS=128 DN_T3_LX=16" poisson3d_256_b1

CUDA events around N eager launches through PreparedEnergy (about 5 us of host time per call),
input sets rotated so that every launch streams from HBM.  A probe tool, not the bench.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workloads", nargs="*", default=["poisson3d_256_b1", "poisson3d_128_b1", "poisson3d_param_64_b16"])
    ap.add_argument("--cfg", action="append", default=[])
    ap.add_argument("--n", type=int, default=40)
    ap.add_argument("--out", default=None)
    ap.add_argument("--graph", action="store_true", help="replay the N launches from one CUDA graph (short kernels)")
    a = ap.parse_args()
    cfgs = [""] + a.cfg
    dev = torch.device("cuda", 0)
    peak, _ = bench.measured_peak_gbs()
    rows = []
    for name in a.workloads:
        nsd, size, B, bpd, _ = bench.WORKLOADS[name]
        fem = bench.make_fem(name)
        dof = B * size ** nsd
        nsets = max(2, min(8, int(-(-400e6 // (dof * bpd)))))
        sets = [bench.make_inputs(name, dev, seed=100 + i) for i in range(nsets)]
        preps = [fem.prepare_energy(s["u"], **bench.call_kwargs(s)) for s in sets]
        ref = None
        for cfg in cfgs:
            saved = {}
            for kv in cfg.split():
                k, v = kv.split("=", 1)
                saved[k] = os.environ.get(k)
                os.environ[k] = v
            try:
                for p in preps:
                    p()
                torch.cuda.synchronize()
                loss, grad = preps[0]()
                l0, g0 = float(loss), grad.clone()
                if ref is None:
                    ref = (l0, g0)
                dl = abs(l0 - ref[0]) / max(abs(ref[0]), 1e-30)
                dg = float((g0 - ref[1]).abs().max() / ref[1].abs().max())
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                best = 1e9
                graph = None
                if a.graph:                      # the knobs are read at capture time: one graph per setting
                    side = torch.cuda.Stream()
                    side.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(side):
                        for p in preps:
                            p()
                    torch.cuda.current_stream().wait_stream(side)
                    torch.cuda.synchronize()
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph, stream=side):
                        for i in range(a.n):
                            preps[i % nsets]()
                    graph.replay()
                for rep in range(3):
                    torch.cuda.synchronize()
                    e0.record()
                    if graph is not None:
                        graph.replay()
                    else:
                        for i in range(a.n):
                            preps[i % nsets]()
                    e1.record()
                    torch.cuda.synchronize()
                    best = min(best, e0.elapsed_time(e1) / a.n)
                gd = dof / (best * 1e-3) / 1e9
                row = dict(workload=name, cfg=cfg, us=best * 1e3, gdofs=gd, frac=gd * bpd / peak, dloss=dl, dgrad=dg)
                print(f"{name:24s} [{cfg:48s}] {best * 1e3:8.1f} us {gd:7.1f} GDOF/s frac {gd * bpd / peak:.3f}  dloss {dl:.1e} dgrad {dg:.1e}",
                      flush=True)
            except Exception as e:   # noqa: BLE001
                row = dict(workload=name, cfg=cfg, error=str(e).splitlines()[0])
                print(f"{name:24s} [{cfg:48s}] FAILED {row['error']}", flush=True)
            rows.append(row)
            for k, v in saved.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
        del sets, preps
        torch.cuda.empty_cache()
    if a.out:
        with open(a.out, "w") as fh:
            for r in rows:
                fh.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
