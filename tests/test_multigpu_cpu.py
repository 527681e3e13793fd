"""N > 1 paths on the CPU (gloo, world_size 2): the host logic of SURVEY.md 8e.

  * z-slab decomposition (diffnet_b200/slab.py): partitioning, halo exchange, ownership of the
    energy and of the gradient, global normalisation, local Adam -- against the whole-domain oracle;
  * data-parallel training (diffnet_b200/trainer.py): DDP gradient averaging over batch shards
    == one process on the whole batch.
The compute backend injected here is the oracle (tests may use it); the product default is the
CUDA op, covered by tests/test_gpu_parity_3d.py::test_z_slab_ownership_matches_whole_domain.
"""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import losses as OL                      # noqa: E402
from oracle.fem import Q1Oracle                      # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _init(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)


def _spawn(fn, world, *args):
    mp.spawn(fn, args=(world, _free_port()) + args, nprocs=world, join=True)


# ------------------------------------------------------------------------------------ z-slabs
NZ, NY, NX = 13, 6, 8


def _global_problem():
    g = torch.Generator().manual_seed(7)
    u = torch.randn(NZ, NY, NX, generator=g, dtype=torch.float64)
    nu = torch.exp(0.4 * torch.randn(NZ, NY, NX, generator=g, dtype=torch.float64))
    f = torch.randn(NZ, NY, NX, generator=g, dtype=torch.float64)
    bc = (torch.rand(NZ, NY, NX, generator=g) > 0.8).double()
    return u, nu, f, bc


def _oracle_backend(lengths):
    """energy(geom, u, z_own=, mean_count=, ...) with the C-ABI's slab semantics, on the oracle:
    loss = energy of the OWNED element layers / mean_count; gradient of ALL local layers."""
    def energy(geom, u, z_own, mean_count, nu=None, f=None, dirichlet=(), **consts):
        o = Q1Oracle(nsd=3, domain_sizes=(geom.nx, geom.ny, geom.nz),
                     domain_lengths=(geom.hx * (geom.nx - 1), geom.hy * (geom.ny - 1), geom.hz * (geom.nz - 1)),
                     dtype=torch.float64)
        b = lambda t: None if t is None else t.reshape((1, 1) + tuple(t.shape[-3:]))
        d = [(b(m), (b(v) if torch.is_tensor(v) else v)) for m, v in dirichlet]
        with torch.enable_grad():
            ul = b(u).detach().clone().requires_grad_(True)
            res = OL.energy_density(o, ul, nu=b(nu), f=b(f), dirichlet=d, **consts)      # (1, ez, ey, ex)
            (grad,) = torch.autograd.grad(res.sum() / mean_count, ul)
        lo, hi = z_own
        hi = min(hi, geom.nz - 1)
        return (res[:, lo:hi].sum() / mean_count).detach(), grad.reshape(u.shape)
    return energy


def _slab_worker(rank, world, port, steps):
    from diffnet_b200 import ops
    from diffnet_b200.slab import ZSlabPoisson3D
    _init(rank, world, port)
    try:
        u, nu, f, bc = _global_problem()
        hx, hy, hz = 1.0 / (NX - 1), 0.8 / (NY - 1), 0.5 / (NZ - 1)
        geom = ops.Geometry(3, NX, NY, NZ, hx, hy, hz, 2)
        sp = ZSlabPoisson3D(geom, energy=_oracle_backend(None))
        sp.set_fields(nu=nu, f=f, dirichlet=[(bc, 0.25)], c_k=0.5)
        # whole-domain truth (every rank computes it: tiny)
        o = Q1Oracle(nsd=3, domain_sizes=(NX, NY, NZ), domain_lengths=(1.0, 0.8, 0.5), dtype=torch.float64)
        b = lambda t: t[None, None]
        ug = u.clone().requires_grad_(True)
        lref = OL.energy_loss(o, b(ug), nu=b(nu), f=b(f), dirichlet=[(b(bc), 0.25)], c_k=0.5)
        (gref,) = torch.autograd.grad(lref, ug)
        # local slab with POISONED halos: the exchange must overwrite them
        ul = sp.local_of(u)
        o0, o1 = sp.slab.own_local
        ul[:o0] = float("nan"); ul[o1:] = float("nan")
        loss, grad = sp.loss_and_grad(ul)
        assert torch.isfinite(ul).all()
        lref = lref.detach()
        assert abs(float(loss) - float(lref)) <= 1e-12 * abs(float(lref)), (float(loss), float(lref))
        gall = sp.gather_owned(grad)
        assert gall.shape == gref.shape
        assert torch.allclose(gall, gref, rtol=1e-11, atol=1e-13)
        assert float(grad[:o0].abs().sum()) == 0.0 and float(grad[o1:].abs().sum()) == 0.0
        # overlapped variant (interior while the halos are in flight, then the boundary pieces)
        ul2 = sp.local_of(u)
        ul2[:o0] = float("nan"); ul2[o1:] = float("nan")
        loss2, grad2 = sp.loss_and_grad(ul2, overlap=True)
        assert abs(float(loss2) - float(lref)) <= 1e-12 * abs(float(lref))
        assert torch.allclose(sp.gather_owned(grad2), gref, rtol=1e-11, atol=1e-13)
        assert float(grad2[:o0].abs().sum()) == 0.0 and float(grad2[o1:].abs().sum()) == 0.0
        # a few Adam steps on the slab == the same steps on the whole domain
        uw = u.clone().requires_grad_(True)
        optw = torch.optim.Adam([uw], lr=0.1)
        ul = sp.local_of(u).requires_grad_(True)
        optl = torch.optim.Adam([ul], lr=0.1)
        for _ in range(steps):
            optw.zero_grad()
            OL.energy_loss(o, b(uw), nu=b(nu), f=b(f), dirichlet=[(b(bc), 0.25)], c_k=0.5).backward()
            optw.step()
            with torch.no_grad():
                _, gl = sp.loss_and_grad(ul)
            ul.grad = gl
            optl.step()
        got = sp.gather_owned(ul.detach())
        assert torch.allclose(got, uw.detach(), rtol=1e-9, atol=1e-11)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_zslab_matches_whole_domain(world):
    _spawn(_slab_worker, world, 3)


def test_slab_bounds_cover_the_domain():
    from diffnet_b200.slab import make_slab, slab_bounds
    for nz in (2, 5, 17, 256):
        for world in (1, 2, 3, 8):
            if nz < world:
                with pytest.raises(ValueError):
                    [make_slab(nz, world, r) for r in range(world)]
                continue
            b = [slab_bounds(nz, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == nz
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            assert max(z1 - z0 for z0, z1 in b) - min(z1 - z0 for z0, z1 in b) <= 1
            for r in range(world):
                s = make_slab(nz, world, r)
                assert s.lo == max(s.z0 - 1, 0) and s.hi == min(s.z1 + 1, nz)


# ------------------------------------------------------------------------------------ DDP
def _make_module():
    from diffnet_b200.base import PDE

    class TinyPoisson(PDE):
        """PoissonParametric2D with the oracle as the loss backend (CPU test double)."""

        def __init__(self, network, **kw):
            super().__init__(network, **kw)
            self.o = Q1Oracle(nsd=2, domain_size=kw["domain_size"])

        def loss(self, u, inputs_tensor, forcing_tensor):
            nu, bc1, bc2 = inputs_tensor[:, 0:1], inputs_tensor[:, 1:2], inputs_tensor[:, 2:3]
            return OL.energy_loss(self.o, u, nu=nu, f=forcing_tensor, dirichlet=[(bc1, 1.0), (bc2, 0.0)])

    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3, padding=1), torch.nn.Tanh(), torch.nn.Conv2d(4, 1, 3, padding=1))
    return TinyPoisson(net, domain_size=12, learning_rate=1e-2)


def _batches(n):
    g = torch.Generator().manual_seed(3)
    out = []
    for _ in range(n):
        nu = torch.exp(0.3 * torch.randn(4, 1, 12, 12, generator=g))
        bc1 = torch.zeros(4, 1, 12, 12); bc1[..., 0] = 1
        bc2 = torch.zeros(4, 1, 12, 12); bc2[..., -1] = 1
        out.append((torch.cat([nu, bc1, bc2], 1), torch.randn(4, 1, 12, 12, generator=g)))
    return out


def _ddp_worker(rank, world, port, ref_state):
    from diffnet_b200.trainer import Trainer, shard_batch
    _init(rank, world, port)
    try:
        m = _make_module()
        tr = Trainer(max_steps=3, device=torch.device("cpu"))
        assert tr.ddp and tr.world == world
        tr.fit(m, [shard_batch(b, rank, world) for b in _batches(3)])
        net = m.network.module
        for k, v in net.state_dict().items():
            assert torch.allclose(v, ref_state[k], rtol=2e-5, atol=1e-7), k
    finally:
        dist.destroy_process_group()


def test_ddp_training_equals_single_process():
    from diffnet_b200.trainer import Trainer
    m = _make_module()
    Trainer(max_steps=3, device=torch.device("cpu"), ddp=False).fit(m, _batches(3))
    ref = {k: v.clone() for k, v in m.network.state_dict().items()}
    _spawn(_ddp_worker, 2, ref)
