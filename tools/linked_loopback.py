"""tools/linked_loopback.py -- the LINKED z-slab launch (dn_fem_energy_3d_linked_f32) on ONE GPU.

A middle rank of a slab chain is linked to ITSELF: its put CTAs store the first / last owned plane into its
own staging planes (first owned -> the "above" halo, last owned -> the "below" halo: a periodic wrap) and
release its own flag words; the CTAs that touch a halo plane wait for those flags inside the same launch
(the put CTAs are the first CTAs of the grid and wait for nothing, so this cannot deadlock -- the ranks are
emulated as one kernel, never as concurrent launches).  Expected result: the plain launch on a slab whose
halo planes hold the wrapped planes.  Used by tests/test_gpu_parity_3d.py and, being one process with no
cross-process waits, it is the form of the linked kernel that can be profiled with ncu.

    python tools/linked_loopback.py [nz ny nx] [--n N]      # times plain vs linked, checks parity
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from diffnet_b200 import _lib as L
from diffnet_b200 import ops


class Loopback:
    """Buffers + dn_slab_link of a rank linked to itself (both neighbours = this rank)."""

    def __init__(self, nl, ny, nx, dev, parity_count=2):
        self.nl, self.ny, self.nx, self.dev = nl, ny, nx, dev
        self.staged = torch.zeros(parity_count, 2, ny * nx, device=dev)          # [parity][below, above]
        self.flags = torch.zeros(parity_count, 2, 32, dtype=torch.int32, device=dev)
        self.ctrl = torch.zeros(parity_count, 8, dtype=torch.int32, device=dev)  # step, status, tickets[2]
        self.slots = torch.zeros(parity_count, 4, dtype=torch.float64, device=dev)   # double[1] + int32 flags
        self.table = [torch.tensor([self.slots[p].data_ptr()], dtype=torch.int64, device=dev)
                      for p in range(parity_count)]

    def link(self, parity=0):
        lk = L.dn_slab_link()
        o0, o1 = 1, self.nl - 1
        st, fl = self.staged[parity], self.flags[parity]
        # halo BELOW (side 0) <- my LAST owned plane (put side 1); halo ABOVE (side 1) <- my FIRST owned plane (put side 0)
        lk.halo_plane[0], lk.halo_flag[0] = st[0].data_ptr(), fl[0].data_ptr()
        lk.halo_plane[1], lk.halo_flag[1] = st[1].data_ptr(), fl[1].data_ptr()
        lk.put_dst[0], lk.put_flag[0], lk.put_plane[0] = st[1].data_ptr(), fl[1].data_ptr(), o0
        lk.put_dst[1], lk.put_flag[1], lk.put_plane[1] = st[0].data_ptr(), fl[0].data_ptr(), o1 - 1
        c = self.ctrl[parity]
        lk.loss_slots = self.table[parity].data_ptr()
        lk.step, lk.status, lk.tickets = c.data_ptr(), c.data_ptr() + 4, c.data_ptr() + 8
        lk.max_spins = 1 << 22
        lk.rank, lk.world = 0, 1
        return lk


def make_case(nl, ny, nx, dev, seed=0):
    g = torch.Generator(device=dev).manual_seed(seed)
    u = torch.randn(nl, ny, nx, device=dev, generator=g)
    nu = torch.rand(nl, ny, nx, device=dev, generator=g) + 0.5
    f = torch.randn(nl, ny, nx, device=dev, generator=g)
    mask = (torch.rand(nl, ny, nx, device=dev, generator=g) > 0.8).float()
    return u, nu, f, mask


def run(nl, ny, nx, dev, n=0, seed=0):
    """(loss_plain, grad_plain, loss_linked, grad_linked, status[, us_plain, us_linked])."""
    h = 1.0 / (nx - 1)
    geom = ops.Geometry(3, nx, ny, nl, h, h, h, 2)
    u, nu, f, mask = make_case(nl, ny, nx, dev, seed)
    kw = dict(nu=nu, f=f, dirichlet=[(mask, 0.0)], c_k=0.5, z_own=(1, nl - 1), mean_count=float((nx - 1) * (ny - 1) * (nl - 2)))
    lb = Loopback(nl, ny, nx, dev)
    linked = ops.PreparedEnergy(geom, u, link=lb.link(0), **kw)
    ll, gl = linked()
    ll, gl = ll.clone(), gl.clone()
    uw = u.clone()
    uw[0], uw[nl - 1] = u[nl - 2], u[1]                       # the periodic wrap the loopback produces
    plain = ops.PreparedEnergy(geom, uw, **kw)
    lp, gp = plain()
    lp, gp = lp.clone(), gp.clone()
    torch.cuda.synchronize()
    out = [lp, gp, ll, gl, int(lb.ctrl[:, 1].sum().item())]
    if n:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for call in (plain, linked):
            for _ in range(5):
                call()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(n):
                call()
            e1.record()
            torch.cuda.synchronize()
            out.append(e0.elapsed_time(e1) / n * 1e3)
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("shape", nargs="*", type=int, default=[34, 256, 256])
    ap.add_argument("--n", type=int, default=50)
    a = ap.parse_args()
    nl, ny, nx = a.shape
    dev = torch.device("cuda", 0)
    lp, gp, ll, gl, status, tp, tl = run(nl, ny, nx, dev, a.n)
    o = slice(1, nl - 1)
    print(f"slab {nl}x{ny}x{nx}: plain {tp:.1f} us  linked {tl:.1f} us  loss rel diff {abs(float(lp) - float(ll)) / abs(float(lp)):.2e}  "
          f"owned grad equal {torch.equal(gp.reshape(nl, ny, nx)[o], gl.reshape(nl, ny, nx)[o])}  wait time-outs {status}")
