"""Launch-shape planners of the streaming kernels (host logic, no GPU): for a sweep of mesh and
batch sizes the chosen shape must be launchable (threads, shared memory, box limits), must cover
the mesh, and its grid must fit the workspace dn_fem_workspace_bytes() promises."""
import ctypes as C
import itertools

import pytest

from diffnet_b200 import _lib as L


def plan(nsd, B, nx, ny, nz, nf=5, has_nu=1):
    g = L.dn_geom(nsd, B, nx, ny, nz, 2, 1.0 / (nx - 1), 1.0 / (ny - 1), 1.0 / max(nz - 1, 1), 0, 0, 0.0)
    out = (C.c_int64 * 16)()
    lib = L.lib()
    assert lib.dn_debug_plan(C.byref(g), nf, has_nu, out) == 0
    return list(out), lib.dn_fem_workspace_bytes(C.byref(g))


@pytest.mark.parametrize("B", [1, 3, 16, 64, 257, 1024])
def test_2d_plans_are_launchable(B):
    for nx, ny in itertools.product([8, 12, 64, 132, 256, 260, 512, 1024, 2048], [2, 3, 9, 64, 255, 256, 513]):
        for nf in (1, 3, 5, 7):
            p, ws = plan(2, B, nx, ny, 1, nf)
            assert p[0] == 1, (B, nx, ny)
            threads, grid, smem, S, R, nch, bal_q, bal_rem = p[1:9]
            assert threads % 32 == 0 and nx // 4 <= threads <= 512
            assert smem <= 226 * 1024 and S >= 2
            if bal_q:       # balanced one-wave split: bal_rem images in bal_q + 1 equal parts, the rest in bal_q
                assert 0 <= bal_rem < B and grid == bal_rem * (bal_q + 1) + (B - bal_rem) * bal_q <= 148 * 8
                for n in {bal_q, bal_q + 1 if bal_rem else bal_q}:
                    cuts = [ch * ny // n for ch in range(n + 1)]
                    if ny % 2 == 0:          # interior cuts at odd rows: every chunk streams whole two-row stages
                        cuts = [c | 1 if 0 < i < n else c for i, c in enumerate(cuts)]
                    assert cuts[0] == 0 and cuts[-1] == ny and min(b - a for a, b in zip(cuts, cuts[1:])) >= 7
                    assert max(b - a for a, b in zip(cuts, cuts[1:])) <= R + 1 <= 33
            else:
                assert R >= 1 and nch * R >= ny and (nch - 1) * R < ny and grid == B * nch
            assert 64 + 8 * grid <= ws
    assert plan(2, B, 2052, 64, 1)[0][0] == 0        # wider than a CTA: the general kernel takes it
    assert plan(2, B, 130, 64, 1)[0][0] == 0         # nx % 4 != 0


@pytest.mark.parametrize("B", [1, 2, 16, 33])
def test_3d_plans_are_launchable(B):
    sizes = [8, 12, 20, 64, 72, 128, 132, 256, 260, 512]
    for nx in sizes:
        for ny, nz in [(2, 2), (5, 3), (16, 16), (64, 64), (130, 40), (256, 256)]:
            if B * nx * ny * nz > 40_000_000:
                continue
            for nf, has_nu in ((1, 0), (4, 0), (5, 1), (7, 1)):
                p, ws = plan(3, B, nx, ny, nz, nf, has_nu)
                assert p[0] == 1, (B, nx, ny, nz)
                threads, grid, smem, S, TY, nty, LXT, ntx, ZC, nzc, BX, BY, rows, LXo, hl = p[1:16]
                assert threads % 32 == 0 and rows * LXT <= threads <= (512 if has_nu else 640)
                assert smem <= 226 * 1024 and S >= 2
                assert nty * TY >= ny and ntx * LXo >= nx // 2 and nzc * ZC >= nz
                assert hl == (1 if ntx > 1 else 0) and LXT == LXo + hl
                assert BX <= 256 and BY <= 256 and BX % 4 == 0 and BX >= 2 * LXT + 2 and BY >= TY + 2 - (1 if nty == 1 else 0)
                assert rows >= min(TY + 1, ny - 1) or nty <= 2
                assert grid == B * nty * ntx * nzc and 64 + 8 * grid <= ws, (B, nx, ny, nz, p)
    assert plan(3, 1, 130, 64, 64)[0][0] == 0
