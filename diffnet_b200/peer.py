"""Halo planes over NVLink peer memory (one process per GPU, CUDA IPC) -- the transport of
``slab.ZSlabPoisson3D(transport="peer")``.

Each rank owns one receive allocation (``dn_peer_alloc``: a raw cudaMalloc, because an IPC handle
names a whole allocation): staging planes ``[parity][side][ny*nx]`` and flag words
``[parity][side]``; its IPC handle is exchanged once (``all_gather_object``) and the two
neighbours map it (``dn_peer_import``, opened on the ACCESSING device).  A step is then four stream-ordered launches of our own kernels
(``dn_peer_put_f32`` x2: stores over NVLink + release of the step counter in the neighbour's flag;
``dn_peer_wait_f32`` x2: bounded device-side wait + copy into the local halo plane) -- no NCCL
call, no host synchronisation, and capturable in a CUDA graph.  Double-buffering by step parity
makes the write-after-read hazards impossible.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _lib as L

BELOW, ABOVE = 0, 1


class PeerHalo:
    def __init__(self, slab, ny: int, nx: int, device: torch.device, group=None, max_spins: int = 1 << 27):
        """``max_spins``: poll limit of every device-side wait (64-ns sleeps between polls: the default is
        on the order of 10 s).  A wait that runs out sets a status word; ``check()`` raises on it."""
        if (ny * nx) % 4:
            raise L.DiffNetFEMError("peer halo planes need ny*nx % 4 == 0")
        self.slab, self.plane, self.device, self.group = slab, ny * nx, device, group
        self.max_spins = max_spins
        self.parity = 0
        plane = self.plane
        self._nbytes = 2 * 2 * plane * 4
        world = dist.get_world_size(group)
        self.world = world
        # after the planes and the 4 halo flags (256 B): two all-reduce areas (parity), each double[world] + int32[world]
        self._red_off = self._nbytes + 256
        self._red_area = (8 * world + 4 * world + 15) // 16 * 16
        self._red_off2 = self._red_off + 2 * self._red_area          # loss slots of the linked steps
        lib = L.lib()
        # local control words per (parity, side): [send counter, put ticket, expect, -, status, wait ticket, -, -]
        self.ctrl = torch.zeros(2, 2, 8, dtype=torch.int32, device=device)
        self._recv = C.c_void_p()
        self._imported = {}
        with torch.cuda.device(device):
            L.check(lib.dn_peer_alloc(self._red_off + 4 * self._red_area, C.byref(self._recv)), "dn_peer_alloc")
            handle = C.create_string_buffer(64)
            L.check(lib.dn_peer_export(self._recv, handle), "dn_peer_export")
        torch.cuda.synchronize(device)
        handles = [None] * world
        dist.all_gather_object(handles, bytes(handle.raw), group=group)
        self._all = {}                                     # rank -> mapped base pointer (every other rank)
        for r in range(world):
            if r != slab.rank:
                ptr = C.c_void_p()
                with torch.cuda.device(device):
                    L.check(lib.dn_peer_import(handles[r], C.byref(ptr)), "dn_peer_import")
                self._all[r] = ptr
        for side, r in ((BELOW, slab.rank - 1), (ABOVE, slab.rank + 1)):
            if 0 <= r < world:
                self._imported[side] = self._all[r]
        # device tables of the peers' all-reduce areas, one per parity; control words for the reduction
        self._red_tables = []
        for par in range(2):
            ptrs = [(self._all[r].value if r != slab.rank else self._recv.value) + self._red_off + par * self._red_area
                    for r in range(world)]
            self._red_tables.append(torch.tensor(ptrs, dtype=torch.int64, device=device))
        self._red_tables2 = []
        for par in range(2):
            ptrs = [(self._all[r].value if r != slab.rank else self._recv.value) + self._red_off2 + par * self._red_area
                    for r in range(world)]
            self._red_tables2.append(torch.tensor(ptrs, dtype=torch.int64, device=device))
        self.red_ctrl = torch.zeros(2, 4, dtype=torch.int32, device=device)       # per parity: [counter, status, -, -]
        self.red_parity = 0
        # linked (one-launch) steps: per parity [step counter, status, ticket0, ticket1]; separate staging flags
        self.link_ctrl = torch.zeros(2, 4, dtype=torch.int32, device=device)
        self.link_parity = 0
        dist.barrier(group=group)                          # everyone mapped before anyone writes

    def close(self):
        lib = L.lib()
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            for ptr in self._all.values():
                lib.dn_peer_unimport(ptr)
            self._all, self._imported = {}, {}
            if self._recv:
                lib.dn_peer_free(self._recv)
                self._recv = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:   # noqa: BLE001  (interpreter shutdown)
            pass

    def _local_ptrs(self, side: int, parity: int):
        slot = parity * 2 + side
        return self._recv.value + slot * self.plane * 4, self._recv.value + self._nbytes + slot * 4

    def _peer_ptrs(self, side: int, parity: int):
        """(staging plane, flag word) in the neighbour on `side`, where I am its OTHER side."""
        base = self._imported[side].value
        slot = parity * 2 + (1 - side)
        return base + slot * self.plane * 4, base + self._nbytes + slot * 4

    def exchange(self, u_local: torch.Tensor) -> None:
        """Refresh the halo planes of the contiguous (nl, ny, nx) slab `u_local` in place."""
        s = self.slab
        o0, o1 = s.own_local
        nl = u_local.shape[0]
        p = self.parity
        self.parity ^= 1
        lib = L.lib()
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        u = u_local.detach()
        esz = 4 * self.plane
        with torch.cuda.device(self.device):
            for side, has, src_plane in ((BELOW, s.has_below, o0), (ABOVE, s.has_above, o1 - 1)):
                if not has:
                    continue
                dst, flag = self._peer_ptrs(side, p)
                c = self.ctrl[p, side]
                L.check(lib.dn_peer_put_f32(C.c_void_p(dst), C.c_void_p(u.data_ptr() + src_plane * esz), self.plane,
                                            C.c_void_p(flag), C.c_void_p(c.data_ptr()), C.c_void_p(c.data_ptr() + 4),
                                            stream), "dn_peer_put_f32")
            for side, has, halo_plane in ((BELOW, s.has_below, 0), (ABOVE, s.has_above, nl - 1)):
                if not has:
                    continue
                c = self.ctrl[p, side]
                staged, flag = self._local_ptrs(side, p)
                L.check(lib.dn_peer_wait_f32(C.c_void_p(u.data_ptr() + halo_plane * esz),
                                             C.c_void_p(staged), self.plane,
                                             C.c_void_p(flag), C.c_void_p(c.data_ptr() + 8),
                                             self.max_spins, C.c_void_p(c.data_ptr() + 16), stream), "dn_peer_wait_f32")

    def allreduce_sum(self, partial: torch.Tensor) -> torch.Tensor:
        """Sum of a 0-dim fp32 device tensor over all ranks through peer memory (rank order, so the
        result is bit-identical everywhere); stream-ordered, graph-capturable, no NCCL."""
        p = self.red_parity
        self.red_parity ^= 1
        out = torch.empty_like(partial)
        lib = L.lib()
        c = self.red_ctrl[p]
        with torch.cuda.device(self.device):
            L.check(lib.dn_peer_allreduce_f32(
                C.c_void_p(partial.data_ptr()), C.c_void_p(out.data_ptr()),
                C.c_void_p(self._recv.value + self._red_off + p * self._red_area),
                C.c_void_p(self._red_tables[p].data_ptr()), self.slab.rank, self.world, C.c_void_p(c.data_ptr()),
                self.max_spins, C.c_void_p(c.data_ptr() + 4),
                C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)), "dn_peer_allreduce_f32")
        return out

    # ------------------------------------------------------------------ linked (one-launch) steps
    def link(self, u_local: torch.Tensor, parity: int) -> "L.dn_slab_link":
        """dn_slab_link of one step with the given parity: the FEM launch itself puts this rank's boundary
        planes, waits for the neighbours' (only in the CTAs that touch a halo), and pushes the loss.
        The linked steps use their own flag words (offset 128 B in the flag block) and step counters, so
        they can be mixed with the put/wait launches of ``exchange``."""
        s = self.slab
        o0, o1 = s.own_local
        lk = L.dn_slab_link()
        esz = 4 * self.plane
        for side, has, src_plane in ((BELOW, s.has_below, o0), (ABOVE, s.has_above, o1 - 1)):
            if has:
                staged, flag = self._local_ptrs(side, parity)
                lk.halo_plane[side] = staged
                lk.halo_flag[side] = flag + 128
                dst, rflag = self._peer_ptrs(side, parity)
                lk.put_dst[side] = dst
                lk.put_flag[side] = rflag + 128
                lk.put_plane[side] = src_plane
        c = self.link_ctrl[parity]
        lk.loss_slots = self._red_tables2[parity].data_ptr()
        lk.step = c.data_ptr()
        lk.status = c.data_ptr() + 4
        lk.tickets = c.data_ptr() + 8
        lk.max_spins = self.max_spins
        lk.rank, lk.world = s.rank, self.world
        return lk

    def loss_sum(self, parity: int) -> torch.Tensor:
        """Global loss of the last linked step of `parity` (sum of the ranks' slots in rank order)."""
        out = torch.empty((), dtype=torch.float32, device=self.device)
        c = self.link_ctrl[parity]
        with torch.cuda.device(self.device):
            L.check(L.lib().dn_peer_loss_sum_f32(
                C.c_void_p(self._recv.value + self._red_off2 + parity * self._red_area), self.world,
                C.c_void_p(c.data_ptr()), self.max_spins, C.c_void_p(c.data_ptr() + 4), C.c_void_p(out.data_ptr()),
                C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)), "dn_peer_loss_sum_f32")
        return out

    def timed_out(self) -> bool:
        """True if any device-side wait hit its poll limit (synchronises)."""
        return (bool(self.ctrl[:, :, 4].any().item()) or bool(self.red_ctrl[:, 1].any().item())
                or bool(self.link_ctrl[:, 1].any().item()))

    def check(self) -> None:
        """Raise if a device-side wait ran out of polls since the last check: the halos / loss of that
        step were stale.  Synchronises the device; call it wherever the host synchronises anyway."""
        if self.timed_out():
            self.ctrl[:, :, 4].zero_(); self.red_ctrl[:, 1].zero_(); self.link_ctrl[:, 1].zero_()
            raise L.DiffNetFEMError(
                "peer halo transport: a device-side wait for a neighbour timed out (rank skew beyond "
                f"max_spins={self.max_spins} polls, or a lost rank); the step that hit it used stale data")
