"""2+ GPU check of the peer-memory halo transport against NCCL send/recv (run under torchrun):
identical loss and gradient, poisoned halos refreshed, then timings eager / graph."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from diffnet_b200 import ops
from diffnet_b200.slab import ZSlabPoisson3D, make_slab

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
h = 1.0 / (N - 1)
geom = ops.Geometry(3, N, N, N, h, h, h, 2)
sl = make_slab(N, world, rank); nl = sl.hi - sl.lo
g = torch.Generator(device=dev).manual_seed(100 + sl.lo)
nu = torch.rand(nl, N, N, device=dev, generator=g) + 0.5
f = torch.ones_like(nu)
bc = (nu > 1.4).float()
# a consistent GLOBAL field: every rank builds the same planes for its stored range
def planes(a, b):
    out = torch.empty(b - a, N, N, device=dev)
    for i, z in enumerate(range(a, b)):
        gz = torch.Generator(device=dev).manual_seed(7000 + z)
        out[i] = torch.randn(N, N, device=dev, generator=gz)
    return out
res = {}
outs = {}
for transport in ("nccl", "peer"):
    sp = ZSlabPoisson3D(geom, transport=transport)
    sp.set_fields(nu=nu, f=f, dirichlet=[(bc, 0.0)], already_local=True, c_k=0.5)
    for it in range(3):                                   # several steps: parity alternation, counters
        u = planes(sl.lo, sl.hi)
        o0, o1 = sp.slab.own_local
        u[:o0] = float("nan"); u[o1:] = float("nan")     # halos must come from the neighbours
        loss, grad = sp.loss_and_grad(u)
        torch.cuda.synchronize()
        assert torch.isfinite(u).all(), f"{transport}: halo not refreshed (step {it})"
    outs[transport] = (loss.clone(), grad.clone(), u.clone())
    if transport == "peer":
        assert not sp._peer_halo.timed_out(), "device-side wait timed out"
    def timeit(fn, n=40, w=8):
        for _ in range(w): fn()
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        for _ in range(n): fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n * 1e3
    res[f"{transport}_eager_ms"] = timeit(lambda: sp.loss_and_grad(u, zero_halo_grad=False))
    res[f"{transport}_exchange_only_ms"] = timeit(lambda: sp.exchange_halos(u))
    if transport == "peer":
        rp = sp.capture(u, zero_halo_grad=False)
        res["peer_graph_ms"] = timeit(rp)
        rp2 = sp.capture(u, zero_halo_grad=False, reduce_loss=False)
        res["peer_graph_noallreduce_ms"] = timeit(rp2)
        lg, gg = rp2()
        lg, gg = rp2()
        lg = sp._reduce(lg, u)
        torch.cuda.synchronize()
        le, ge = sp.loss_and_grad(u, zero_halo_grad=False)
        le2, _ = sp.loss_and_grad(u, zero_halo_grad=False)
        ln_ = le.clone(); dist.all_reduce(ln_)                # NCCL sum of the (already global) value
        res["peer_allreduce_vs_nccl_rel"] = abs(float(ln_) / world - float(le)) / abs(float(le))
        res["graph_vs_eager_loss_rel"] = abs(float(lg) - float(le)) / abs(float(le))
        res["graph_vs_eager_grad_max"] = float((gg - ge)[sp.slab.own_local[0]:sp.slab.own_local[1]].abs().max())
        assert not sp._peer_halo.timed_out(), "device-side wait timed out (graph)"
# ---- linked (one-launch) steps: same loss (after the on-demand sum) and gradient, bit for bit
sp = ZSlabPoisson3D(geom, transport="peer")
sp.set_fields(nu=nu, f=f, dirichlet=[(bc, 0.0)], already_local=True, c_k=0.5)
uref = outs["peer"][2]                                   # halos refreshed by the put/wait path
o0, o1 = sp.slab.own_local
for it in range(3):
    ul = uref.clone()
    ul[:o0] = 12345.0; ul[o1:] = -54321.0                # the linked kernel must NOT read the local halo planes
    lpart, gl = sp.step_linked(ul)
    ltot = sp.global_loss()
    torch.cuda.synchronize()
sp.check()
res["linked_loss_equal"] = bool(torch.equal(ltot, outs["peer"][0]))
res["linked_loss_rel"] = abs(float(ltot) - float(outs["peer"][0])) / abs(float(outs["peer"][0]))
res["linked_grad_equal"] = bool(torch.equal(gl[o0:o1], outs["peer"][1][o0:o1]))
res["linked_eager_ms"] = timeit(lambda: sp.step_linked(uref))
rpl = sp.capture(uref, linked=True)
res["linked_graph_ms"] = timeit(rpl)
lg2, gg2 = rpl(); lg2, gg2 = rpl()
ltot2 = sp.global_loss()
torch.cuda.synchronize()
sp.check()
res["linked_graph_loss_equal"] = bool(torch.equal(ltot2, outs["peer"][0]))
res["linked_graph_grad_equal"] = bool(torch.equal(gg2[o0:o1], outs["peer"][1][o0:o1]))
ln, gn, un = outs["nccl"]; lp, gp, up = outs["peer"]
res["loss_equal"] = bool(torch.equal(ln, lp)); res["grad_equal"] = bool(torch.equal(gn, gp)); res["u_equal"] = bool(torch.equal(un, up))
if rank == 0:
    print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in res.items()}, flush=True)
dist.destroy_process_group()
