"""tools/gp_probe.py [n] -- the un-fused gauss_pt_evaluation forward / multi-table / adjoint launches at the two
BASELINE batch shapes (256^2 x 64, 64^3 x 16), timed with CUDA events; the command ncu profiles for gp_eval."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from diffnet_b200 import DiffNet2DFEM, DiffNet3DFEM, ops

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dev = "cuda:0"
peak = 6455.6
for fem, shape in ((DiffNet2DFEM(None, domain_size=256), (64, 1, 256, 256)), (DiffNet3DFEM(None, domain_size=64), (16, 1, 64, 64, 64))):
    us = [torch.randn(shape, device=dev) for _ in range(4)]
    ngp = fem.ngp_total
    nel = 1
    for s in fem.geometry.elems:
        nel *= s
    nodes = us[0].numel()
    outs = ops._gp_raw(fem.geometry, us[0], 0)
    cots = [torch.randn_like(outs) for _ in range(2)]
    which = (0, 1, 2) if fem.nsd == 2 else (0, 1, 2, 3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, nbytes, label):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        us_ = e0.elapsed_time(e1) / n * 1e3
        print(f"{fem.nsd}-D {label:28s} {us_:8.1f} us  {nbytes / us_ / 1e3:7.0f} GB/s  {nbytes / us_ / 1e3 / peak:.3f} of peak", flush=True)

    u_req = [u.clone().requires_grad_(True) for u in us]
    timed(lambda i: ops._gp_raw(fem.geometry, us[i % 4], 0), 4 * nodes + 4 * shape[0] * ngp * nel, "forward, one table")
    timed(lambda i: ops._gp_multi_raw(fem.geometry, us[i % 4], which), 4 * nodes + 4 * len(which) * shape[0] * ngp * nel, f"forward, {len(which)} tables, one pass")

    held = [ops.gp_eval(fem.geometry, u, "N") for u in u_req[:2]]       # the adjoint alone: grad of a held output

    def adj(i):
        torch.autograd.grad(held[i % 2], u_req[i % 2], cots[i % 2], retain_graph=True)
    timed(adj, 4 * nodes + 4 * shape[0] * ngp * nel, "adjoint (autograd.grad)")

    heldm = [ops.gp_eval_multi(fem.geometry, u, ("N", "dx", "dy", "dz")[:len(which)]) for u in u_req[:2]]     # all tables: one adjoint call
    cotm = [tuple(torch.randn_like(o) for o in heldm[0]) for _ in range(2)]

    def adjm(i):
        torch.autograd.grad(heldm[i % 2], u_req[i % 2], cotm[i % 2], retain_graph=True)
    timed(adjm, 4 * nodes + 4 * len(which) * shape[0] * ngp * nel, f"adjoint, {len(which)} tables, one call")


# device-side kernel durations (the eager numbers above include ~30 us of host / autograd time per call)
from torch.profiler import ProfilerActivity, profile

for fem, shape in ((DiffNet2DFEM(None, domain_size=256), (64, 1, 256, 256)), (DiffNet3DFEM(None, domain_size=64), (16, 1, 64, 64, 64))):
    u = torch.randn(shape, device=dev, requires_grad=True)
    out = ops.gp_eval(fem.geometry, u, "N")
    cot = torch.randn_like(out)
    torch.autograd.grad(out, u, cot, retain_graph=True)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(10):
            ops._gp_raw(fem.geometry, u.detach(), 0)
            torch.autograd.grad(out, u, cot, retain_graph=True)
        torch.cuda.synchronize()
    nel = 1
    for s_ in fem.geometry.elems:
        nel *= s_
    nbytes = 4 * u.numel() + 4 * shape[0] * fem.ngp_total * nel
    for ev in prof.key_averages():
        if "k_gp_eval" in ev.key:
            t = ev.device_time_total / ev.count
            print(f"{fem.nsd}-D kernel {ev.key[:40]:40s} {t:7.1f} us  {nbytes / t / 1e3:6.0f} GB/s  {nbytes / t / 1e3 / peak:.3f} of peak", flush=True)
