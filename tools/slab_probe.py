"""Timing probe for the pieces of one z-slab step (run under torchrun on >= 2 GPUs)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from diffnet_b200 import ops
from diffnet_b200.slab import ZSlabPoisson3D, make_slab

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
N = 256; h = 1.0 / (N - 1)
geom = ops.Geometry(3, N, N, N, h, h, h, 2)
sp = ZSlabPoisson3D(geom); sl = make_slab(N, world, rank); nl = sl.hi - sl.lo
u = torch.randn(nl, N, N, device=dev)
nu = torch.rand(nl, N, N, device=dev) + 0.5
sp.set_fields(nu=nu, f=torch.ones_like(nu), dirichlet=[((nu > 1.4).float(), 0.0)], already_local=True, c_k=0.5)

def timeit(fn, n=30, w=5):
    for _ in range(w): fn()
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3

o0, o1 = sp.slab.own_local
def ex_staged():
    opsl, keep = [], []
    if sl.has_below:
        s_ = u[o0].contiguous(); r_ = torch.empty_like(s_); keep.append((r_, 0))
        opsl += [dist.P2POp(dist.isend, s_, rank - 1), dist.P2POp(dist.irecv, r_, rank - 1)]
    if sl.has_above:
        s_ = u[o1 - 1].contiguous(); r_ = torch.empty_like(s_); keep.append((r_, nl - 1))
        opsl += [dist.P2POp(dist.isend, s_, rank + 1), dist.P2POp(dist.irecv, r_, rank + 1)]
    for w_ in dist.batch_isend_irecv(opsl): w_.wait()
    for r_, pl in keep: u[pl].copy_(r_)
sbuf = torch.empty(2, N, N, device=dev); rbuf = torch.empty(2, N, N, device=dev)
def ex_persist():
    opsl = []
    if sl.has_below:
        sbuf[0].copy_(u[o0]); opsl += [dist.P2POp(dist.isend, sbuf[0], rank - 1), dist.P2POp(dist.irecv, rbuf[0], rank - 1)]
    if sl.has_above:
        sbuf[1].copy_(u[o1 - 1]); opsl += [dist.P2POp(dist.isend, sbuf[1], rank + 1), dist.P2POp(dist.irecv, rbuf[1], rank + 1)]
    for w_ in dist.batch_isend_irecv(opsl): w_.wait()
    if sl.has_below: u[0].copy_(rbuf[0])
    if sl.has_above: u[nl - 1].copy_(rbuf[1])
res = {}
res["exchange_inplace_ms"] = timeit(lambda: sp.exchange_halos(u))
res["exchange_staged_ms"] = timeit(ex_staged)
res["exchange_persistent_ms"] = timeit(ex_persist)
res["kernel_only_ms"] = timeit(lambda: sp.loss_and_grad(u, exchange=False, zero_halo_grad=False, reduce_loss=False))
l = torch.zeros((), device=dev)
res["allreduce_scalar_ms"] = timeit(lambda: dist.all_reduce(l))
res["full_step_ms"] = timeit(lambda: sp.loss_and_grad(u, zero_halo_grad=False))
res["full_step_overlap_ms"] = timeit(lambda: sp.loss_and_grad(u, zero_halo_grad=False, overlap=True))
la, ga = sp.loss_and_grad(u.clone(), zero_halo_grad=True); lb, gb = sp.loss_and_grad(u.clone(), zero_halo_grad=True, overlap=True)
res["overlap_vs_plain_loss_rel"] = abs(float(la) - float(lb)) / abs(float(la)); res["overlap_vs_plain_grad_maxabs"] = float((ga - gb).abs().max())
if rank == 0:
    print({k: round(v, 4) for k, v in res.items()}, flush=True)
dist.destroy_process_group()
