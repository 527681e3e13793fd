"""3-D parity: the CUDA path (through the C ABI) vs. the oracle and the golden vectors."""
import os

import numpy as np
import pytest
import torch

from conftest import rel_l2, rel_scalar
from helpers import GRAD_RTOL, LOSS_RTOL, assert_parity, oracle_energy, oracle_residual
from diffnet_b200 import DiffNet3DFEM

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def make_inputs(B, D, H, W, seed=0):
    g = torch.Generator().manual_seed(seed)
    u = torch.randn(B, 1, D, H, W, generator=g)
    nu = torch.exp(0.5 * torch.randn(B, 1, D, H, W, generator=g))
    f = torch.randn(B, 1, D, H, W, generator=g)
    src = (torch.rand(B, 1, D, H, W, generator=g) > 0.93).float()
    sink = torch.zeros(B, 1, D, H, W)
    sink[:, :, 0] = 1; sink[:, :, -1] = 1; sink[:, :, :, 0] = 1
    sink[:, :, :, -1] = 1; sink[..., 0] = 1; sink[..., -1] = 1
    return u, nu, f, src, sink


def dev(x):
    if torch.is_tensor(x):
        return x.to(DEV)
    if isinstance(x, (list, tuple)):
        return type(x)(dev(v) for v in x)
    return x


def run_energy(fem, u, **kw):
    ud = u.to(DEV).requires_grad_(True)
    loss = fem.energy_loss(ud, **{k: dev(v) for k, v in kw.items()})
    loss.backward()
    return loss.detach().cpu(), ud.grad.detach().cpu()


SIZES = [(1, 5, 6, 8), (2, 5, 6, 7), (1, 9, 11, 13), (2, 12, 20, 16), (1, 20, 33, 64), (1, 10, 12, 128),
         (1, 7, 9, 132), (1, 6, 10, 260), (1, 5, 8, 520), (2, 16, 16, 16)]


@pytest.mark.parametrize("B,D,H,W", SIZES)
def test_energy_sizes(B, D, H, W):
    """Full Poisson (u, nu, f, two masks) over aligned / odd / multi-tile meshes."""
    fem = DiffNet3DFEM(None, domain_sizes=(W, H, D), domain_lengths=(1.0, 0.8, 0.5), domain_size=W)
    u, nu, f, src, sink = make_inputs(B, D, H, W, seed=D * 10000 + H * 100 + W)
    kw = dict(nu=nu, f=f, dirichlet=[(sink, 0.0), (src, 1.0)])
    loss, grad = run_energy(fem, u, **kw)
    lref, gref = oracle_energy(fem, u, **kw)
    assert_parity(loss, grad, lref, gref, masks=(src, sink), what=f"3D {B}x{D}x{H}x{W}")


@pytest.mark.parametrize("variant", ["ibn3d", "inobj", "nomask", "numask", "E6", "sum"])
def test_energy_family(variant):
    B, D, H, W = 2, 14, 18, 24
    fem = DiffNet3DFEM(None, domain_sizes=(W, H, D), domain_size=W)
    u, nu, f, src, sink = make_inputs(B, D, H, W, seed=17)
    kw = {
        "ibn3d": dict(f=f, dirichlet=[(sink, 0.0), (src, 1.0)]),                 # IBN_3D.py:114-136
        "inobj": dict(nu=nu, f=f, dirichlet=[(src, 0.0)], c_k=0.5),              # solve_in_object_3d.py:75-102
        "nomask": dict(nu=nu, f=f),
        "numask": dict(nu=nu, f=f, nu_zero_mask=src, dirichlet=[(sink, 0.0)]),
        "E6": dict(nu=nu, dirichlet=[(sink, 0.0), (src, 1.0)], c_k=0.5, c_f=0.0),  # 9_voxel_3d.py:119
        "sum": dict(nu=nu, f=f, dirichlet=[(sink, 0.0)], reduction="sum", scale=0.125),
    }[variant]
    loss, grad = run_energy(fem, u, **kw)
    lref, gref = oracle_energy(fem, u, **kw)
    assert_parity(loss, grad, lref, gref, masks=[m for m, _ in kw.get("dirichlet", [])], what=variant)


@pytest.mark.parametrize("ngp", [2, 3])
def test_f_at_gauss_points_and_value_field(ngp):
    """e8_3d_poisson_mms.py:48,143,165: forcing at Gauss points + nodal Dirichlet field."""
    B, D, H, W = 1, 8, 10, 12
    fem = DiffNet3DFEM(None, domain_sizes=(W, H, D), domain_size=W, ngp_1d=ngp)
    u, nu, f, src, sink = make_inputs(B, D, H, W, seed=23)
    u_bc = torch.randn(1, 1, D, H, W)
    f_gp = torch.randn(1, ngp ** 3, D - 1, H - 1, W - 1)
    kw = dict(nu=nu, f_gp=f_gp, dirichlet=[(sink, u_bc)], c_k=0.5)
    loss, grad = run_energy(fem, u, **kw)
    lref, gref = oracle_energy(fem, u, **kw)
    assert_parity(loss, grad, lref, gref, masks=(sink,), what=f"fgp ngp{ngp}")


def test_golden_vectors(golden):
    g = golden("ref_3d_box")
    X, Y, Z = (int(v) for v in g["sizes"])
    fem = DiffNet3DFEM(None, domain_sizes=(X, Y, Z), domain_lengths=(1.0, 0.8, 0.5), domain_size=X)
    u, inputs, f = g.t("u"), g.t("inputs"), g.t("forcing")
    src, sink = g.t("source"), g.t("sink")
    loss, grad = run_energy(fem, u, f=f, dirichlet=[(sink, 0.0), (src, 1.0)])
    assert rel_scalar(loss, g["ibn3d.loss64"]) <= LOSS_RTOL and rel_l2(grad, g["ibn3d.grad64"]) <= GRAD_RTOL
    loss, grad = run_energy(fem, u, nu=inputs[:, 0:1], f=f, dirichlet=[(inputs[:, 1:2], 0.0)], c_k=0.5)
    assert rel_scalar(loss, g["inobj.loss64"]) <= LOSS_RTOL and rel_l2(grad, g["inobj.grad64"]) <= GRAD_RTOL
    # bare (D,H,W) parameter, B = 1 (solve_in_object_3d.py:198-199)
    ub = u[0, 0].to(DEV).requires_grad_(True)
    inp = inputs[:1].to(DEV)
    loss = fem.energy_loss(ub, nu=inp[:, 0:1], f=f[:1].to(DEV), dirichlet=[(inp[:, 1:2], 0.0)], c_k=0.5)
    loss.backward()
    assert tuple(ub.grad.shape) == (Z, Y, X)
    assert rel_scalar(loss.cpu(), g["inobj_bare.loss64"]) <= LOSS_RTOL
    assert rel_l2(ub.grad.cpu(), g["inobj_bare.grad64"]) <= GRAD_RTOL
    for key, fn in (("N", fem.gauss_pt_evaluation), ("dx", fem.gauss_pt_evaluation_der_x),
                    ("dy", fem.gauss_pt_evaluation_der_y), ("dz", fem.gauss_pt_evaluation_der_z)):
        out = fn(u.to(DEV))
        assert out.shape == g["gp." + key].shape
        assert rel_l2(out.cpu(), g["gp." + key]) <= 1e-6, key


def test_golden_forcing_at_gauss_points(golden):
    """The real reference's 3-D module on the e8_3d_poisson_mms.py form (tests/golden/make_golden_fgp3d.py): Dirichlet
    value field + f at the Gauss points, here through the assembled load vector of the streaming kernel (nx = 8)."""
    g = golden("ref_3d_fgp")
    X, Y, Z = (int(v) for v in g["sizes"])
    fem = DiffNet3DFEM(None, domain_sizes=(X, Y, Z), domain_lengths=(1.0, 0.8, 0.5), domain_size=X)
    loss, grad = run_energy(fem, g.t("u"), nu=g.t("nu"), f_gp=g.t("f_gp"), dirichlet=[(g.t("bc"), g.t("u_bc"))], c_k=0.5)
    assert rel_scalar(loss, g["E3fgp.loss64"]) <= LOSS_RTOL and rel_l2(grad, g["E3fgp.grad64"]) <= GRAD_RTOL
    hit = (g.t("bc") > 0.5).expand_as(grad)
    assert torch.count_nonzero(grad[hit]) == 0


def test_residual_form():
    B, D, H, W = 1, 9, 10, 12
    fem = DiffNet3DFEM(None, domain_sizes=(W, H, D), domain_size=W)
    u, nu, f, src, sink = make_inputs(B, D, H, W, seed=29)
    d = [(sink, 0.0), (src, 1.0)]
    ud = u.to(DEV).requires_grad_(True)
    loss = fem.residual_loss(ud, nu=nu.to(DEV), f=f.to(DEV), dirichlet=dev(d), jac=(0.5 * fem.h) ** 3)
    loss.backward()
    lref, gref = oracle_residual(fem, u, nu=nu, f=f, dirichlet=d, jac=(0.5 * fem.h) ** 3)
    assert_parity(loss.cpu(), ud.grad.cpu(), lref, gref, masks=(src, sink), what="resmin3d")


def test_unfused_body_on_gp_eval_ops():
    B, D, H, W = 1, 7, 8, 9
    fem = DiffNet3DFEM(None, domain_sizes=(W, H, D), domain_size=W)
    u, nu, f, _, _ = make_inputs(B, D, H, W, seed=31)
    ud = u.to(DEV).requires_grad_(True)
    gsq = (fem.gauss_pt_evaluation_der_x(ud) ** 2 + fem.gauss_pt_evaluation_der_y(ud) ** 2
           + fem.gauss_pt_evaluation_der_z(ud) ** 2)
    res = fem.gauss_pt_evaluation(nu.to(DEV)) * gsq - fem.gauss_pt_evaluation(ud) * fem.gauss_pt_evaluation(f.to(DEV))
    loss = torch.mean(torch.sum(res, 1))
    loss.backward()
    lref, gref = oracle_energy(fem, u, nu=nu, f=f)
    assert_parity(loss.cpu(), ud.grad.cpu(), lref, gref, what="unfused 3d")


def test_full_size_properties():
    """64^3 B=16, 128^3 and 256^3 B=1 (BASELINE sizes): Euler identity, constants, scaling, determinism,
    independence of the z-chunking / tile shape."""
    for B, N in ((16, 64), (1, 128), (1, 256)):
        fem = DiffNet3DFEM(None, domain_size=N)
        g = torch.Generator(device=DEV).manual_seed(N)
        u = torch.randn(B, 1, N, N, N, device=DEV, generator=g)
        nu = torch.exp(0.3 * torch.randn(B, 1, N, N, N, device=DEV, generator=g))
        loss, grad = fem.energy_loss_and_grad(u, nu=nu)
        euler = float((grad.double() * u[:, 0].double()).sum())
        assert rel_scalar(euler, 2.0 * float(loss)) < 2e-5
        l2, g2 = fem.energy_loss_and_grad(2.0 * u, nu=nu)
        assert rel_scalar(l2, 4.0 * float(loss)) < 1e-5 and rel_l2(g2, 2.0 * grad) < 1e-5
        lc, gc = fem.energy_loss_and_grad(torch.full_like(u, 3.0), nu=nu)
        assert float(lc) == 0.0 and float(gc.abs().max()) == 0.0
        la, ga = fem.energy_loss_and_grad(u, nu=nu)
        assert torch.equal(la, loss) and torch.equal(ga, grad)
        for knobs in ({"DN_ZC_3D": "11", "DN_T3_ZC": "11"}, {"DN_ROWS_3D": "6", "DN_T3_TY": "5"},
                      {"DN_3D_PATH": "tile"}):
            os.environ.update(knobs)
            try:
                lb, gb = fem.energy_loss_and_grad(u, nu=nu)
            finally:
                for k in knobs:
                    os.environ.pop(k)
            assert rel_scalar(lb, loss) < 2e-6 and rel_l2(gb, grad) < 2e-6, knobs


@pytest.mark.parametrize("N", [64, 128])
def test_baseline_grids_against_oracle(N):
    """64^3 and 128^3 (B = 1 so the fp64 oracle finishes in seconds) through the streaming
    kernel, full Poisson (u, nu, f, two masks)."""
    fem = DiffNet3DFEM(None, domain_size=N)
    u, nu, f, src, sink = make_inputs(1, N, N, N, seed=N)
    kw = dict(nu=nu, f=f, dirichlet=[(sink, 0.0), (src, 1.0)])
    loss, grad = run_energy(fem, u, **kw)
    lref, gref = oracle_energy(fem, u, **kw)
    assert_parity(loss, grad, lref, gref, masks=(src, sink), what=f"3D {N}^3")


def test_bench_launch_shape_64cubed_batch16_against_oracle():
    """configs[2] as the bench launches it (64^3, B = 16, nu == 1, source/sink masks): one sample of the
    full-batch launch against the fp64 oracle on that sample."""
    from diffnet_b200.synthetic import poisson3d_parametric_batch
    B, N, b = 16, 64, 11
    fem = DiffNet3DFEM(None, domain_size=N, batch_size=B)
    u, src, sink, f = poisson3d_parametric_batch(B, N, DEV, seed=5)
    f = f + torch.randn_like(f)
    loss, grad = fem.energy_loss_and_grad(u, f=f, dirichlet=[(sink, 0.0), (src, 1.0)], reduction="sum")
    sl = slice(b, b + 1)
    kwb = dict(f=f[sl], dirichlet=[(sink[sl], 0.0), (src[sl], 1.0)])
    lref, gref = oracle_energy(fem, u[sl], reduction="sum", **kwb)
    lb, gb = fem.energy_loss_and_grad(u[sl], reduction="sum", **kwb)
    assert_parity(lb, grad[b], lref, gref[0, 0], masks=(sink[b, 0], src[b, 0]), what="sample 11 of the 64^3 x 16 launch")
    assert rel_l2(gb[0], grad[b]) < 2e-6


def test_streaming_and_tile_paths_agree():
    """k_fem3d_tma (bulk-async streaming) vs k_fem3d (general tile kernel): same operator."""
    B, D, H, W = 2, 19, 37, 72
    fem = DiffNet3DFEM(None, domain_sizes=(W, H, D), domain_lengths=(1.0, 0.7, 0.4), domain_size=W)
    u, nu, f, src, sink = (t.to(DEV) for t in make_inputs(B, D, H, W, seed=11))
    ubc = torch.randn_like(u)
    top = torch.zeros_like(src); top[:, :, :, 0, :] = 1
    cases = [dict(), dict(nu=nu), dict(f=f), dict(nu=nu, f=f, dirichlet=[(src, 1.0)]),
             dict(nu=nu, f=f, dirichlet=[(sink, 0.0), (src, 1.0)]),
             dict(f=f, dirichlet=[(sink, 0.0), (src, 1.0), (top, 0.25)]),
             dict(nu=nu, f=f, dirichlet=[(sink, ubc)]),
             dict(nu=nu, f=f, nu_zero_mask=top, dirichlet=[(sink, 0.0), (src, 1.0)])]
    for kw in cases:
        os.environ.pop("DN_3D_PATH", None)
        ls, gs = fem.energy_loss_and_grad(u, **kw)
        os.environ["DN_3D_PATH"] = "tile"
        try:
            lw, gw = fem.energy_loss_and_grad(u, **kw)
        finally:
            os.environ.pop("DN_3D_PATH", None)
        assert rel_scalar(ls, lw) < 2e-6, (sorted(kw), float(ls), float(lw))
        assert rel_l2(gs, gw) < 2e-6, sorted(kw)


def test_residual_form_streaming_vs_general():
    B, D, H, W = 1, 11, 14, 40
    fem = DiffNet3DFEM(None, domain_sizes=(W, H, D), domain_size=W)
    u, nu, f, src, sink = (t.to(DEV) for t in make_inputs(B, D, H, W, seed=21))
    out = {}
    for path in ("", "tile"):
        if path:
            os.environ["DN_3D_PATH"] = path
        try:
            ud = u.clone().requires_grad_(True)
            loss = fem.residual_loss(ud, nu=nu, f=f, dirichlet=[(sink, 0.0), (src, 1.0)], jac=(0.5 * fem.h) ** 3)
            loss.backward()
            out[path] = (loss.detach(), ud.grad.detach())
        finally:
            os.environ.pop("DN_3D_PATH", None)
    assert rel_scalar(out[""][0], out["tile"][0]) < 2e-6
    assert rel_l2(out[""][1], out["tile"][1]) < 2e-6


def test_z_slab_ownership_matches_whole_domain():
    """slabs with one-plane halos, loss summed over owned layers, gradient complete
    on owned planes, divided by the GLOBAL element count -- emulated on one GPU."""
    from diffnet_b200 import ops
    B, N = 1, 24
    fem = DiffNet3DFEM(None, domain_size=N)
    u, nu, f, src, sink = make_inputs(B, N, N, N, seed=37)
    u, nu, f, src = (t.to(DEV) for t in (u, nu, f, src))
    d = [(src, 0.0)]
    loss, grad = fem.energy_loss_and_grad(u, nu=nu, f=f, dirichlet=d, c_k=0.5)
    total, parts = 0.0, []
    nslab = 3
    count = float(B * (N - 1) ** 3)
    for s in range(nslab):
        z0, z1 = s * N // nslab, (s + 1) * N // nslab
        lo, hi = max(z0 - 1, 0), min(z1 + 1, N)
        geom = ops.Geometry(3, N, N, hi - lo, fem.hx, fem.hy, fem.hz, 2)
        sl = lambda t: t[:, :, lo:hi]
        l, g, _ = ops.energy_raw(geom, sl(u), nu=sl(nu), f=sl(f), dirichlet=[(sl(src), 0.0)], c_k=0.5,
                                 z_own=(z0 - lo, z1 - lo), mean_count=count)
        total += float(l)
        parts.append(g[:, z0 - lo:z1 - lo])
    assert rel_scalar(total, float(loss)) < 1e-5
    assert rel_l2(torch.cat(parts, 1), grad) < 1e-6


def test_linked_slab_launch_on_one_gpu():
    """dn_fem_energy_3d_linked_f32 (halo exchange + loss push inside the FEM launch) with the
    neighbours EMULATED by pre-filled staging planes and released flags, so nothing waits: the halo
    planes are read from staging (the slab's own halo planes hold junk), the boundary planes land
    in the 'neighbours'' buffers with the launch number in their flags, the loss partial lands in
    the slot table, and loss / owned gradient equal the plain launch on the slab with halos in place."""
    import ctypes as C
    from diffnet_b200 import _lib as L, ops
    N, lo, hi = 32, 9, 23                      # stored planes [lo, hi): owned [lo+1, hi-1)
    fem = DiffNet3DFEM(None, domain_size=N)
    u, nu, f, src, sink = (t.to(DEV)[0, 0] for t in make_inputs(1, N, N, N, seed=91))
    nl, o0, o1 = hi - lo, 1, hi - lo - 1
    geom = ops.Geometry(3, N, N, nl, fem.hx, fem.hy, fem.hz, 2)
    count = float((N - 1) ** 3)
    sl = lambda t: t[lo:hi].contiguous()       # noqa: E731
    kw = dict(nu=sl(nu), f=sl(f), dirichlet=[(sl(src), 0.0)], c_k=0.5)
    lref, gref, _ = ops.energy_raw(geom, sl(u), z_own=(o0, o1), mean_count=count, **kw)
    ul = u[lo:hi].clone()                      # (a slice along z is already contiguous: sl() would alias u)
    ul[0] = 777.0
    ul[-1] = -777.0
    staging = torch.stack([u[lo], u[hi - 1]]).contiguous()
    flags = torch.zeros(64, dtype=torch.int32, device=DEV)          # [0,1] halo flags, [32,33] "remote" flags
    remote = torch.full((2, N * N), float("nan"), device=DEV)
    area = torch.zeros(4, dtype=torch.float64, device=DEV)          # double[1] slot + int32[1] flag (+ slack)
    table = torch.tensor([area.data_ptr()], dtype=torch.int64, device=DEV)
    ctrl = torch.zeros(4, dtype=torch.int32, device=DEV)            # step, status, tickets[2]
    lk = L.dn_slab_link()
    for sd, plane in ((0, o0), (1, o1 - 1)):
        lk.halo_plane[sd] = staging[sd].data_ptr()
        lk.halo_flag[sd] = flags.data_ptr() + 4 * sd
        lk.put_dst[sd] = remote[sd].data_ptr()
        lk.put_flag[sd] = flags.data_ptr() + 4 * (32 + sd)
        lk.put_plane[sd] = plane
    lk.loss_slots = table.data_ptr()
    lk.step, lk.status, lk.tickets = ctrl.data_ptr(), ctrl.data_ptr() + 4, ctrl.data_ptr() + 8
    lk.max_spins, lk.rank, lk.world = 1 << 16, 0, 1
    call = ops.PreparedEnergy(geom, ul, z_own=(o0, o1), mean_count=count, link=lk, **kw)
    for step in (1, 2, 3):
        flags[:2] = step                        # the neighbours have "released" this step
        remote.fill_(float("nan"))
        loss, grad = call()
        torch.cuda.synchronize()
        assert int(ctrl[0]) == step and int(ctrl[1]) == 0
        assert torch.equal(grad.reshape(nl, N, N)[o0:o1], gref.reshape(nl, N, N)[o0:o1])
        assert torch.equal(loss, lref)
        assert torch.equal(remote[0], ul[o0].reshape(-1)) and torch.equal(remote[1], ul[o1 - 1].reshape(-1))
        assert flags[32:34].tolist() == [step, step]
        assert float(area[0]) == pytest.approx(float(lref), rel=1e-7)
        assert int(area[1:2].view(torch.int32)[0]) == step
    # a flag that never arrives: bounded wait, status word set, no hang
    flags[:2] = 0
    call()
    torch.cuda.synchronize()
    assert int(ctrl[1]) == 1


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs on one node")
def test_peer_memory_halo_transport_two_gpus():
    """tools/slab_peer_check.py under torchrun on 2 GPUs: the NVLink peer-memory transport
    (dn_peer_put/wait) gives bit-identical loss, gradient and refreshed halos to NCCL send/recv,
    eagerly and replayed from CUDA graphs."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533",
                        os.path.join(root, "tools", "slab_peer_check.py"), "64"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "'loss_equal': True" in r.stdout and "'grad_equal': True" in r.stdout and "'u_equal': True" in r.stdout
    assert "'linked_grad_equal': True" in r.stdout and "'linked_graph_grad_equal': True" in r.stdout


def test_randomised_parity_sweep():
    """tools/fuzz_parity.py for a few seconds: random sizes / views / Dirichlet sets / options, 2-D and
    3-D, streaming kernels vs the general kernels and (small cases) the fp64 oracle.  (A 240 s run
    of the same script covered 34,569 cases without a failure.)"""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "fuzz_parity.py"), "8", "7"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "cases ok" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]


@pytest.mark.parametrize("B,D,H,W,ngp", [(2, 9, 21, 70, 2), (1, 5, 12, 33, 2), (1, 4, 6, 9, 3)])
def test_gp_eval_marching_kernels_against_oracle(B, D, H, W, ngp):
    """3-D marching gp-eval kernels + the multi-table pass (N, dx, dy, dz) against the oracle's convs in fp64."""
    from helpers import oracle_for
    fem = DiffNet3DFEM(None, domain_sizes=(W, H, D), domain_lengths=(1.3, 0.9, 0.6), domain_size=W, ngp_1d=ngp)
    o = oracle_for(fem)
    g = torch.Generator().manual_seed(B * 1000 + W)
    u = torch.randn(B, 1, D, H, W, generator=g)
    ud = u.to(DEV).requires_grad_(True)
    outs = fem.gauss_pt_evaluation_all(ud)
    uo = u.double().requires_grad_(True)
    refs = (o.gauss_pt_evaluation(uo), o.gauss_pt_evaluation_der_x(uo), o.gauss_pt_evaluation_der_y(uo),
            o.gauss_pt_evaluation_der_z(uo))
    cot = [torch.randn(r.shape, generator=g) for r in refs]
    for a, r in zip(outs, refs):
        assert a.shape == r.shape
        assert rel_l2(a.detach().cpu(), r.detach()) <= 1e-6
    assert torch.equal(outs[3], fem.gauss_pt_evaluation_der_z(ud))
    sum((a * c.to(DEV)).sum() for a, c in zip(outs, cot)).backward()
    sum((r * c.double()).sum() for r, c in zip(refs, cot)).backward()
    assert rel_l2(ud.grad.cpu(), uo.grad) <= 1e-6
    assert float((ud.grad.cpu().double() - uo.grad).abs().max() / uo.grad.abs().max()) <= 2e-6


@pytest.mark.parametrize("B,D,H,W,ngp", [(1, 2, 2, 2, 2), (3, 8, 8, 32, 2), (2, 9, 8, 33, 2), (1, 20, 17, 130, 2),
                                         (2, 17, 15, 64, 2), (1, 10, 10, 40, 3), (1, 9, 9, 34, 4)])
def test_gp_eval_adjoint_z_march_against_oracle_and_y_march(B, D, H, W, ngp, monkeypatch):
    """The z-marching 3-D adjoint (k_gp_eval_adj3: one load per element value, shared-memory x/y exchange, register
    z carry) over tile-exact, one-past-a-tile and multi-tile / multi-chunk sizes: against the transpose of the
    oracle's conv in fp64, and against the y-marching kernel (DN_GP_ADJ3=0) it replaces."""
    from helpers import oracle_for
    fem = DiffNet3DFEM(None, domain_sizes=(W, H, D), domain_lengths=(1.3, 0.9, 0.6), domain_size=W, ngp_1d=ngp)
    o = oracle_for(fem)
    g = torch.Generator().manual_seed(D * 100 + W)
    u = torch.randn(B, 1, D, H, W, generator=g)
    for name in ("gauss_pt_evaluation", "gauss_pt_evaluation_der_x", "gauss_pt_evaluation_der_z"):
        uo = u.double().requires_grad_(True)
        ref = getattr(o, name)(uo)
        cot = torch.randn(ref.shape, generator=g)
        (gref,) = torch.autograd.grad(ref, uo, cot.double())
        got = []
        for flag in ("1", "0"):
            monkeypatch.setenv("DN_GP_ADJ3", flag)
            ud = u.to(DEV).requires_grad_(True)
            (gd,) = torch.autograd.grad(getattr(fem, name)(ud), ud, cot.to(DEV))
            got.append(gd.cpu())
            assert float((gd.cpu().double() - gref).abs().max() / gref.abs().max()) <= 2e-6, (name, flag)
        assert rel_l2(got[0], got[1]) <= 1e-6


@pytest.mark.parametrize("shape", [(6, 16, 32), (10, 40, 72), (34, 64, 64)])
def test_linked_slab_launch_loopback_one_gpu(shape):
    """dn_fem_energy_3d_linked_f32 on ONE GPU: a middle rank linked to itself (tools/linked_loopback.py) --
    in-launch halo put, flag wait, downward march of the lower chunk and the loss push -- against the plain
    launch on the slab with the wrapped halo planes: owned gradient bit for bit, loss to rounding."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    from linked_loopback import run
    nl, ny, nx = shape
    lp, gp, ll, gl, status = run(nl, ny, nx, torch.device(DEV), seed=nl)
    assert status == 0
    assert rel_scalar(ll, lp) < 2e-6
    o = slice(1, nl - 1)
    assert torch.equal(gl.reshape(nl, ny, nx)[o], gp.reshape(nl, ny, nx)[o])


@pytest.mark.parametrize("ngp,numask", [(2, False), (3, False), (2, True)])
def test_grad_nu_3d(ngp, numask):
    """d loss / d nu in 3-D (the fused launch for loss + dL/du, the gather launch k_grad_nu_3d for dL/dnu):
    Dirichlet masks applied to u, optional nu mask, grad_output scaling -- against oracle autograd in fp64."""
    from helpers import oracle_for, to64
    from oracle import losses as OL
    B, D, H, W = 2, 7, 9, 12
    fem = DiffNet3DFEM(None, domain_sizes=(W, H, D), domain_lengths=(1.0, 0.8, 0.5), domain_size=W, ngp_1d=ngp)
    u, nu, f, src, sink = make_inputs(B, D, H, W, seed=41)
    d = [(sink, 0.0), (src, 1.0)]
    kw = dict(nu_zero_mask=src) if numask else {}
    ud = u.to(DEV).requires_grad_(True)
    nud = nu.to(DEV).clone().requires_grad_(True)
    loss = 2.5 * fem.energy_loss(ud, nu=nud, f=f.to(DEV), dirichlet=dev(d), c_k=0.5,
                                 **{k: v.to(DEV) for k, v in kw.items()})
    loss.backward()
    o = oracle_for(fem)
    u64, nu64 = to64(u).requires_grad_(True), to64(nu).requires_grad_(True)
    lref = 2.5 * OL.energy_loss(o, u64, nu=nu64, f=to64(f), dirichlet=to64(d), c_k=0.5,
                                **{k: to64(v) for k, v in kw.items()})
    gu, gn = torch.autograd.grad(lref, (u64, nu64))
    assert rel_scalar(loss.cpu(), lref) <= LOSS_RTOL
    assert rel_l2(ud.grad.cpu(), gu) <= GRAD_RTOL
    assert rel_l2(nud.grad.cpu(), gn) <= GRAD_RTOL
    assert float((nud.grad.cpu().double() - gn).abs().max() / gn.abs().max()) <= GRAD_RTOL
    if numask:
        assert float(nud.grad.cpu()[src > 0.5].abs().max()) == 0.0
