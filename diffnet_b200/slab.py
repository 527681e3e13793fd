"""z-slab domain decomposition of ONE large 3-D field over the ranks of a process group
(non-parametric case; new functionality -- the reference runs 128^3 on one GPU,
IBN/poisson-3d/non-parametric/solve_in_object_3d.py).

Rank k owns node planes [z0_k, z1_k) and stores planes [z0_k - 1, z1_k + 1) clipped to the
domain: a one-plane halo on each side.  Static fields (nu, f, masks) are cut once with their
halos.  Every step:
  1. halo exchange of u: each rank sends its first / last OWNED plane to the rank below / above
     (2 planes x ny x nx x 4 B -- 256 KB at 256^2) with batched point-to-point ops (NCCL send/recv
     over NVLink on GPUs, gloo on CPU);
  2. the fused kernel runs on the local slab: the energy is summed over element layers whose
     lower plane is owned (``z_own``), divided by the GLOBAL element count (``mean_count``); the
     gradient of the owned planes is complete without a reverse exchange because the kernel
     evaluates the halo element layers too;
  3. one scalar all-reduce(SUM) of the loss partials (only the value: gradients are local);
  4. the optimiser (Adam on u, solve_in_object_3d.py:123) updates the owned planes locally.

The compute backend is a callable ``energy(geom, u, z_own=, mean_count=, **fields) -> (loss, grad)``;
the product default is the CUDA op.  The CPU tests (gloo, world_size 2) inject the oracle there to
check the partitioning, the exchange and the global normalisation without a GPU.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional, Sequence

import torch
import torch.distributed as dist

from . import ops


def slab_bounds(nz: int, world: int, rank: int) -> tuple[int, int]:
    """Owned node planes [z0, z1) of `rank`: contiguous, sizes differing by at most one."""
    base, rem = divmod(nz, world)
    z0 = rank * base + min(rank, rem)
    return z0, z0 + base + (1 if rank < rem else 0)


@dataclass
class Slab:
    rank: int
    world: int
    nz: int                 # global node planes
    z0: int                 # owned [z0, z1)
    z1: int
    lo: int                 # stored [lo, hi) = owned + halos, clipped
    hi: int

    @property
    def own_local(self) -> tuple[int, int]:
        return self.z0 - self.lo, self.z1 - self.lo

    @property
    def has_below(self) -> bool:
        return self.z0 > 0

    @property
    def has_above(self) -> bool:
        return self.z1 < self.nz


def make_slab(nz: int, world: int, rank: int) -> Slab:
    z0, z1 = slab_bounds(nz, world, rank)
    if z1 - z0 < 1:
        raise ValueError(f"rank {rank} owns no plane: nz={nz} < world={world}")
    return Slab(rank, world, nz, z0, z1, max(z0 - 1, 0), min(z1 + 1, nz))


def cut(t: torch.Tensor, slab: Slab, zdim: int = -3) -> torch.Tensor:
    """The stored planes [lo, hi) of a global field (a copy: each rank keeps only its slab)."""
    return t.narrow(zdim, slab.lo, slab.hi - slab.lo).contiguous()


_prepared = {}


def _default_energy(geom, u, **kw):
    """The CUDA op.  Slab steps call it on the same storage every iteration: the marshalled call
    (ops.PreparedEnergy) is cached per (storage, shape, ownership)."""
    key = (u.data_ptr(), tuple(u.shape), geom, kw.get("z_own"), kw.get("mean_count"),
           tuple((m.data_ptr(), v.data_ptr() if torch.is_tensor(v) else v) for m, v in kw.get("dirichlet", ())),
           tuple((k, v.data_ptr()) for k, v in kw.items() if torch.is_tensor(v)),
           tuple((k, v) for k, v in kw.items() if isinstance(v, (int, float))))
    call = _prepared.get(key)
    if call is None:
        if len(_prepared) > 64:
            _prepared.clear()
        call = _prepared[key] = ops.PreparedEnergy(geom, u, **kw)
    loss, grad = call()
    return loss, grad


class ZSlabPoisson3D:
    """Loss + gradient of the Poisson energy of one (nz, ny, nx) field split into z-slabs.

    ``u_local`` is this rank's (hi-lo, ny, nx) slab INCLUDING halos; only the owned planes are
    optimisation variables, the halos are refreshed from the neighbours at every step."""

    def __init__(self, geom_global: ops.Geometry, group=None, energy: Optional[Callable] = None,
                 transport: str = "nccl"):
        if geom_global.nsd != 3:
            raise ValueError("z-slab decomposition is for 3-D meshes")
        self.g = geom_global
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.slab = make_slab(geom_global.nz, self.world, self.rank)
        nl = self.slab.hi - self.slab.lo
        self.geom_local = ops.Geometry(3, geom_global.nx, geom_global.ny, nl, geom_global.hx, geom_global.hy,
                                       geom_global.hz, geom_global.ngp_1d)
        self.energy = energy or _default_energy
        if transport not in ("nccl", "peer"):
            raise ValueError("transport must be 'nccl' (send/recv) or 'peer' (NVLink peer memory, CUDA only)")
        self.transport = transport
        self._peer_halo = None                 # PeerHalo, created at the first exchange (needs the device)
        self._linked = {}                      # prepared linked launches per (storage, parity)
        self._last_linked_parity = 0
        self.fields = {}
        self.dirichlet: Sequence = ()
        self.consts = {}

    # ---------------------------------------------------------------- setup
    def set_fields(self, nu=None, f=None, dirichlet=(), already_local=False, **consts):
        """Static fields: global tensors (cut here, halos included) or, with ``already_local``,
        this rank's slabs.  ``consts`` = c_k, c_f, scale."""
        c = (lambda t: t) if already_local else (lambda t: cut(t, self.slab))
        self.fields = {k: c(v) for k, v in (("nu", nu), ("f", f)) if v is not None}
        self.dirichlet = [(c(m), (c(v) if torch.is_tensor(v) else v)) for m, v in dirichlet]
        self.consts = consts

    def local_of(self, u_global: torch.Tensor) -> torch.Tensor:
        return cut(u_global, self.slab)

    # ---------------------------------------------------------------- per step
    def start_exchange(self, u_local: torch.Tensor):
        """Enqueue the halo exchange (asynchronous w.r.t. the compute stream) and return the
        pending work handles.  Planes of a contiguous (nl, ny, nx) slab are contiguous, so they are
        sent and received in place (no staging copies); the point-to-point ops go out as one NCCL
        group.  NCCL orders the group after everything already on the current stream."""
        if self.world == 1:
            return []
        if not u_local.is_contiguous():
            raise ValueError("u_local must be contiguous (planes are exchanged in place)")
        s = self.slab
        o0, o1 = s.own_local
        u = u_local.detach()
        opsl = []
        if s.has_below:     # my first owned plane -> rank-1's upper halo; its last owned -> my lower halo
            opsl += [dist.P2POp(dist.isend, u[o0], self._peer(-1), self.group),
                     dist.P2POp(dist.irecv, u[0], self._peer(-1), self.group)]
        if s.has_above:
            opsl += [dist.P2POp(dist.isend, u[o1 - 1], self._peer(+1), self.group),
                     dist.P2POp(dist.irecv, u[u.shape[0] - 1], self._peer(+1), self.group)]
        return dist.batch_isend_irecv(opsl)

    def exchange_halos(self, u_local: torch.Tensor) -> None:
        """In place: halo planes of `u_local` <- the neighbours' boundary owned planes."""
        if self.transport == "peer" and self.world > 1:
            if not u_local.is_contiguous():
                raise ValueError("u_local must be contiguous (planes are exchanged in place)")
            if self._peer_halo is None:
                from .peer import PeerHalo
                self._peer_halo = PeerHalo(self.slab, self.g.ny, self.g.nx, u_local.device, self.group)
            self._peer_halo.exchange(u_local)
            return
        for w in self.start_exchange(u_local):
            w.wait()

    def _peer(self, d: int) -> int:
        r = self.rank + d
        return dist.get_global_rank(self.group, r) if self.group is not None else r

    def _sub(self, t, a, b):
        return None if t is None else t[a:b]

    def _energy_on(self, u_local, a, b, z_own, count):
        """The fused kernel on local planes [a, b) (views, no copies)."""
        fields = {k: v[a:b] for k, v in self.fields.items()}
        dirichlet = [(m[a:b], (v[a:b] if torch.is_tensor(v) else v)) for m, v in self.dirichlet]
        g = self.geom_local
        geom = ops.Geometry(3, g.nx, g.ny, b - a, g.hx, g.hy, g.hz, g.ngp_1d)
        return self.energy(geom, u_local[a:b], z_own=z_own, mean_count=count, dirichlet=dirichlet,
                           **fields, **self.consts)

    def loss_and_grad(self, u_local: torch.Tensor, exchange: bool = True, zero_halo_grad: bool = True,
                      reduce_loss: bool = True, overlap: bool = False):
        """(loss, d loss / d u on this rank's slab).  With ``reduce_loss`` the loss is the GLOBAL
        value (one scalar all-reduce), identical on every rank; otherwise this rank's partial.
        The gradient is complete on the owned planes; the halo planes are zeroed unless
        ``zero_halo_grad=False`` (their values are then partial sums nobody should use).

        ``overlap=True`` hides the exchange behind the interior: the owned planes are processed
        while the halo planes are in flight (their energy and the gradient of the planes strictly
        inside), then the two 3-plane boundary pieces close the gradient of the first / last owned
        plane and add the energy of the layer that touches the upper halo."""
        g = self.g
        count = float((g.nx - 1) * (g.ny - 1) * (g.nz - 1))
        s = self.slab
        o0, o1 = s.own_local
        nl = u_local.shape[0]
        if overlap and exchange and self.world > 1 and self.transport == "peer":
            raise NotImplementedError(
                "overlap=True starts the halos with NCCL point-to-point; with transport='peer' use step_linked(), "
                "which overlaps the exchange with the interior planes inside one launch")
        if not (overlap and exchange and self.world > 1 and (o1 - o0) >= 4):
            if exchange:
                self.exchange_halos(u_local)
            loss, grad = self._energy_on(u_local, 0, nl, (o0, o1), count)
            grad = grad.reshape(u_local.shape)
        else:
            pending = self.start_exchange(u_local)
            # interior: owned planes only; every element layer between two owned planes counts
            loss, gi = self._energy_on(u_local, o0, o1, (0, o1 - o0), count)
            grad = torch.empty_like(u_local) if zero_halo_grad is False else torch.zeros_like(u_local)
            grad[o0:o1] = gi.reshape((o1 - o0,) + tuple(u_local.shape[1:]))
            for w in pending:
                w.wait()
            none = (1 << 20, (1 << 20) + 1)           # an ownership range that matches no layer
            if s.has_below:       # planes o0-1, o0, o0+1: closes grad[o0]; its layers are counted elsewhere
                _, gb = self._energy_on(u_local, o0 - 1, o0 + 2, none, count)
                grad[o0] = gb.reshape((3,) + tuple(u_local.shape[1:]))[1]
            if s.has_above:       # planes o1-2, o1-1, o1: closes grad[o1-1]; the layer above o1-1 is ours
                la, ga = self._energy_on(u_local, o1 - 2, o1 + 1, (1, 2), count)
                grad[o1 - 1] = ga.reshape((3,) + tuple(u_local.shape[1:]))[1]
                loss = loss + la
        if zero_halo_grad and not (overlap and exchange and self.world > 1 and (o1 - o0) >= 4):
            if o0 > 0:
                grad[:o0].zero_()
            if o1 < grad.shape[0]:
                grad[o1:].zero_()
        if reduce_loss and self.world > 1:
            loss = self._reduce(loss, u_local)
        return loss, grad

    # ---------------------------------------------------------------- linked (one-launch) step
    def step_linked(self, u_local: torch.Tensor):
        """ONE launch per rank: the FEM kernel itself stores this rank's boundary planes of u into the
        neighbours' staging buffers, waits for theirs only in the CTAs that touch a halo plane (the
        interior planes run meanwhile), and pushes the rank's loss partial to every rank
        (``include/diffnet_fem.h: dn_slab_link``).  Returns (this rank's loss PARTIAL, grad) -- the
        same static tensors on every call with the same storage; ``global_loss()`` adds the partials up
        when the value is wanted.  The halo planes of ``u_local`` itself are NOT refreshed (the kernel
        reads the staged planes), and the gradient of the halo planes is a partial sum nobody should
        use.  Needs transport='peer' and CUDA; with one rank it is the plain fused launch."""
        if self.world == 1:
            return self.loss_and_grad(u_local, exchange=False, zero_halo_grad=False, reduce_loss=False)
        if self.transport != "peer" or not u_local.is_cuda:
            raise NotImplementedError("step_linked needs transport='peer' on CUDA")
        if not u_local.is_contiguous():
            raise ValueError("u_local must be contiguous")
        if self._peer_halo is None:
            from .peer import PeerHalo
            self._peer_halo = PeerHalo(self.slab, self.g.ny, self.g.nx, u_local.device, self.group)
        ph = self._peer_halo
        par = ph.link_parity
        ph.link_parity ^= 1
        key = (u_local.data_ptr(), tuple(u_local.shape), par)
        call = self._linked.get(key)
        if call is None:
            g = self.g
            count = float((g.nx - 1) * (g.ny - 1) * (g.nz - 1))
            o0, o1 = self.slab.own_local
            call = self._linked[key] = ops.PreparedEnergy(
                self.geom_local, u_local, z_own=(o0, o1), mean_count=count, dirichlet=self.dirichlet,
                link=ph.link(u_local, par), **self.fields, **self.consts)
        self._last_linked_parity = par
        loss, grad = call()
        return loss, grad.reshape(u_local.shape)

    def global_loss(self) -> torch.Tensor:
        """Global loss of the most recent ``step_linked`` (sum of all ranks' partials, rank order)."""
        if self.world == 1 or self._peer_halo is None:
            raise RuntimeError("global_loss() follows step_linked() on more than one rank")
        return self._peer_halo.loss_sum(self._last_linked_parity)

    def check(self) -> None:
        """Raise if a device-side wait of the peer transport timed out (synchronises)."""
        if self._peer_halo is not None:
            self._peer_halo.check()

    def _reduce(self, loss, u_local):
        if self.transport == "peer" and loss.is_cuda:
            if self._peer_halo is None:
                from .peer import PeerHalo
                self._peer_halo = PeerHalo(self.slab, self.g.ny, self.g.nx, u_local.device, self.group)
            return self._peer_halo.allreduce_sum(loss.reshape(()).float())
        loss = loss.clone()
        dist.all_reduce(loss, op=dist.ReduceOp.SUM, group=self.group)
        return loss

    def _parity(self):
        """(halo parity, all-reduce parity, linked-step parity) of the peer transport."""
        ph = self._peer_halo
        return (ph.parity, ph.red_parity, ph.link_parity) if ph is not None else (0, 0, 0)

    def _set_parity(self, p):
        ph = self._peer_halo
        if ph is not None:
            ph.parity, ph.red_parity, ph.link_parity = p

    def capture(self, u_locals, warmup: int = 4, linked: bool = False, **kw):
        """Capture whole steps into CUDA graphs and return ``replay() -> (loss, grad)`` (static output
        tensors of the step just replayed).  ``linked=False``: halo put/wait launches + FEM kernel +
        peer all-reduce of the loss; ``linked=True``: the single linked launch of ``step_linked`` (the
        loss returned is then this rank's partial).
        ``u_locals``: one slab tensor or a list of them (the graphs are bound to their storage and
        replayed round-robin: rotating input sets for benchmarks).  At these sizes a step is a few
        host-launched operations of 10-20 us each around a 20-100 us kernel: replaying a graph removes
        the host from the critical path.  Every rank must capture and replay in lockstep.  With
        world > 1 this needs ``transport="peer"`` (our kernels are plain launches; the captured NCCL
        send/recv group hung on this stack: torch 2.11 / NCCL 2.28.9, 2 x B200).  The double-buffering
        parities are part of each graph: graph i uses the parities (start + i) % 2 (an even number of
        graphs is captured), and replay() keeps the host-side parities in step so that eager calls
        may be mixed in between."""
        us = list(u_locals) if isinstance(u_locals, (list, tuple)) else [u_locals]
        if not all(u.is_cuda for u in us):
            raise ValueError("capture() needs CUDA tensors")
        if self.world > 1 and self.transport != "peer":
            raise NotImplementedError("graph capture with world > 1 needs transport='peer'")
        reduce_loss = kw.pop("reduce_loss", True)
        dev = us[0].device

        def one(u):
            if linked:
                return self.step_linked(u)
            return self.loss_and_grad(u, reduce_loss=reduce_loss, **kw)

        ngraphs = len(us) if (self.world == 1 or len(us) % 2 == 0) else 2 * len(us)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for i in range(2 * max(warmup // 2, 1)):      # allocates workspaces, maps the peers; even count
                one(us[i % len(us)])
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        start = self._parity()
        flip = lambda p, i: tuple((q + i) % 2 for q in p)   # noqa: E731
        graphs, outs = [], []
        for i in range(ngraphs):
            self._set_parity(flip(start, i))
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                outs.append(one(us[i % len(us)]))
            graphs.append(g)
            # the capture ran nothing: run the step for real so that every rank's device-side
            # counters advance exactly as the parity sequence says
            g.replay()
        self._set_parity(flip(start, ngraphs))
        torch.cuda.synchronize(dev)
        state = {"i": 0}

        def replay():
            i = state["i"]
            if self.world > 1 and self._peer_halo is not None and self._parity() != flip(start, i):
                raise RuntimeError("graph replay out of step with the halo parity (an odd number of eager "
                                   "steps was mixed in); run one more eager step or re-capture")
            graphs[i].replay()
            state["i"] = (i + 1) % ngraphs
            self._set_parity(flip(start, i + 1))
            if linked:
                self._last_linked_parity = flip(start, i)[2]
            return outs[i]
        replay.graphs = graphs
        return replay

    def gather_owned(self, u_local: torch.Tensor) -> torch.Tensor:
        """All ranks' owned planes concatenated in z (for checks / output)."""
        o0, o1 = self.slab.own_local
        mine = u_local[o0:o1].contiguous()
        self.check()                                 # a host synchronisation point: stale halos must not pass silently
        if self.world == 1:
            return mine
        sizes = [slab_bounds(self.g.nz, self.world, r) for r in range(self.world)]
        most = max(b - a for a, b in sizes)          # all_gather wants equal shapes: pad to the largest slab
        pad = torch.zeros((most,) + tuple(mine.shape[1:]), dtype=mine.dtype, device=mine.device)
        pad[:mine.shape[0]] = mine
        bufs = [torch.empty_like(pad) for _ in sizes]
        dist.all_gather(bufs, pad, group=self.group)
        return torch.cat([b[:hi - lo] for b, (lo, hi) in zip(bufs, sizes)], 0)
