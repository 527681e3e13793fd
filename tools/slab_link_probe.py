"""tools/slab_link_probe.py -- where the time of a linked z-slab step goes (run under torchrun, >= 2 GPUs):
the plain fused launch on this rank's slab (no exchange, no reduction) against the linked launch (puts + flag
waits + loss push inside the launch), both replayed from CUDA graphs, per rank."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from diffnet_b200 import ops
from diffnet_b200.slab import ZSlabPoisson3D, make_slab

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
N = int(os.environ.get("PROBE_N", "256"))
K = 100
h = 1.0 / (N - 1)
NZ = int(os.environ.get("PROBE_NZ", str(N)))          # planes of the whole field (default: a cube)
geom = ops.Geometry(3, N, N, NZ, h, h, h, 2)


def run(tag):
    sp = ZSlabPoisson3D(geom, transport="peer")
    sl = make_slab(NZ, world, rank)
    nl = sl.hi - sl.lo
    g = torch.Generator(device=dev).manual_seed(7 + rank)
    us = [torch.randn(nl, N, N, device=dev, generator=g) for _ in range(4)]
    nu = torch.rand(nl, N, N, device=dev, generator=g) + 0.5
    sp.set_fields(nu=nu, f=torch.full_like(nu, 500.0), dirichlet=[((nu > 1.4).float(), 0.0)], already_local=True, c_k=0.5)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(step):
        for _ in range(8):
            step()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0.record()
        for _ in range(K):
            step()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / K * 1e3

    plain = sp.capture(us, exchange=False, zero_halo_grad=False, reduce_loss=False)
    t_plain = timed(plain)
    linked = sp.capture(us, linked=True)
    t_link = timed(linked)
    t_sep = 0.0
    if not os.environ.get("DN_SLAB_DBG"):
        sep = sp.capture(us, zero_halo_grad=False)
        t_sep = timed(sep)
    if not os.environ.get("DN_SLAB_DBG"):
        sp.check()
    row = torch.tensor([t_plain, t_link, t_sep], device=dev, dtype=torch.float64)
    rows = [torch.zeros_like(row) for _ in range(world)]
    dist.all_gather(rows, row)
    if rank == 0:
        print(json.dumps({"cfg": tag, "world": world, "planes_rank0": nl,
                          "us_per_step [plain kernel, linked step, separate launches] per rank":
                              [[round(float(v), 1) for v in r] for r in rows]}), flush=True)


for cfg in os.environ.get("PROBE_CFGS", ";DN_SLAB_NPUT=8;DN_SLAB_NPUT=1").split(";"):
    saved = {}
    for kv in cfg.split():
        k, v = kv.split("=", 1)
        saved[k] = os.environ.get(k)
        os.environ[k] = v
    try:
        run(cfg)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
dist.destroy_process_group()
