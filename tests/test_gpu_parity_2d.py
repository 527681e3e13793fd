"""2-D parity: the CUDA path (through the C ABI) vs. the oracle and the golden vectors.

Tolerances (north_star): loss rel <= 1e-5, gradient rel-L2 <= 1e-4, gradient exactly 0 on
Dirichlet nodes.  The oracle truth is evaluated in fp64 on the CPU.
"""
import math
import os

import numpy as np
import pytest
import torch

from conftest import rel_l2, rel_scalar
from helpers import GRAD_RTOL, LOSS_RTOL, assert_parity, oracle_energy, oracle_residual
from diffnet_b200 import DiffNet2DFEM

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def make_inputs(B, H, W, seed=0, smooth=False):
    g = torch.Generator().manual_seed(seed)
    if smooth:
        yy, xx = torch.meshgrid(torch.linspace(0, 1, H), torch.linspace(0, 1, W), indexing="ij")
        u = (torch.sin(np.pi * xx) * torch.sin(np.pi * yy))[None, None].repeat(B, 1, 1, 1)
        u = u + 0.01 * torch.randn(B, 1, H, W, generator=g)
    else:
        u = torch.randn(B, 1, H, W, generator=g)
    nu = torch.exp(0.5 * torch.randn(B, 1, H, W, generator=g))
    f = torch.randn(B, 1, H, W, generator=g)
    bc1 = torch.zeros(B, 1, H, W); bc1[..., 0] = 1
    bc2 = torch.zeros(B, 1, H, W); bc2[..., -1] = 1
    inputs = torch.cat([nu, bc1, bc2], 1)
    return u, inputs, f


def run_energy(fem, u, **kw):
    """CUDA loss + grad via autograd (the user-facing path)."""
    ud = u.to(DEV).requires_grad_(True)
    kwd = {}
    for k, v in kw.items():
        if torch.is_tensor(v):
            kwd[k] = v.to(DEV)
        elif k == "dirichlet":
            kwd[k] = [(m.to(DEV), (val.to(DEV) if torch.is_tensor(val) else val)) for m, val in v]
        else:
            kwd[k] = v
    loss = fem.energy_loss(ud, **kwd)
    loss.backward()
    return loss.detach().cpu(), ud.grad.detach().cpu()


SIZES = [(1, 8, 8), (2, 9, 12), (3, 21, 37), (2, 64, 64), (1, 66, 130), (2, 40, 256), (1, 33, 260),
         (2, 128, 512), (1, 17, 516)]


@pytest.mark.parametrize("B,H,W", SIZES)
def test_energy_e1_sizes(B, H, W):
    """E1 (12_klsum.py:53-78) over aligned, odd (scalar path) and multi-strip widths."""
    fem = DiffNet2DFEM(None, domain_sizes=(W, H, 1), domain_size=W)
    u, inputs, f = make_inputs(B, H, W, seed=H * 1000 + W)
    nu, bc1, bc2 = inputs[:, 0:1], inputs[:, 1:2], inputs[:, 2:3]
    kw = dict(nu=nu, f=f, dirichlet=[(bc1, 1.0), (bc2, 0.0)])
    loss, grad = run_energy(fem, u, **kw)
    lref, gref = oracle_energy(fem, u, **kw)
    assert_parity(loss, grad, lref, gref, masks=(bc1, bc2), what=f"E1 {B}x{H}x{W}")


@pytest.mark.parametrize("variant", ["E2", "E3", "E4", "E5", "E6", "nomask", "onemask", "threemask"])
def test_energy_family(variant):
    """The loss family of SURVEY.md App. A.4 on one 48 x 72 mesh."""
    B, H, W = 3, 48, 72
    fem = DiffNet2DFEM(None, domain_sizes=(W, H, 1), domain_lengths=(1.5, 1.0, 1.0), domain_size=W)
    u, inputs, f = make_inputs(B, H, W, seed=7)
    nu, bc1, bc2 = inputs[:, 0:1], inputs[:, 1:2], inputs[:, 2:3]
    obj = torch.zeros_like(nu); obj[:, :, 10:20, 30:50] = 1
    kw = {
        "E2": dict(nu=nu, f=f, dirichlet=[(bc1, 1.0), (bc2, 0.0)], scale=0.5 * (0.5 * fem.h) ** 2),
        "E3": dict(nu=nu, f=f, dirichlet=[(bc1, 1.0), (bc2, 0.0)], c_k=0.5),
        "E4": dict(f=f, dirichlet=[(bc1, 1.0), (bc2, 0.0)]),
        "E5": dict(nu=nu, f=f, nu_zero_mask=obj, dirichlet=[(bc1, 1.0), (bc2, 0.0)]),
        "E6": dict(nu=nu, dirichlet=[(bc1, 1.0), (bc2, 0.0)], c_k=0.5, c_f=0.0),
        "nomask": dict(nu=nu, f=f),
        "onemask": dict(nu=nu, f=f, dirichlet=[(obj, 0.25)], reduction="sum"),
        "threemask": dict(nu=nu, f=f, dirichlet=[(obj, 0.5), (bc1, 1.0), (bc2, 0.0)]),
    }[variant]
    loss, grad = run_energy(fem, u, **kw)
    lref, gref = oracle_energy(fem, u, **kw)
    masks = [m for m, _ in kw.get("dirichlet", [])]
    assert_parity(loss, grad, lref, gref, masks=masks, what=variant)


def test_mask_precedence_and_overlap():
    """Later Dirichlet entries win where masks overlap (0_base.py:41-42); IBN_3D-style
    'source wins over sink' is the reversed order."""
    B, H, W = 2, 24, 32
    fem = DiffNet2DFEM(None, domain_sizes=(W, H, 1), domain_size=W)
    u, inputs, f = make_inputs(B, H, W, seed=3)
    a = torch.zeros(B, 1, H, W); a[:, :, 4:12, 4:20] = 1
    b = torch.zeros(B, 1, H, W); b[:, :, 8:16, 10:28] = 1
    for order in ([(a, 1.0), (b, 0.0)], [(b, 0.0), (a, 1.0)]):
        kw = dict(nu=inputs[:, 0:1], f=f, dirichlet=order)
        loss, grad = run_energy(fem, u, **kw)
        lref, gref = oracle_energy(fem, u, **kw)
        assert_parity(loss, grad, lref, gref, masks=(a, b), what="overlap")


def test_dirichlet_value_field_and_f_at_gauss_points():
    """E3 with u = where(bc, u_bc, u) and analytic forcing at the Gauss points
    (e8_2d_poisson_mms.py:47,154,165,175)."""
    B, H, W = 2, 30, 44
    fem = DiffNet2DFEM(None, domain_sizes=(W, H, 1), domain_size=W)
    u, inputs, _ = make_inputs(B, H, W, seed=11)
    nu = inputs[:, 0:1]
    edge = torch.zeros(1, 1, H, W); edge[..., 0] = 1; edge[..., -1] = 1; edge[:, :, 0] = 1; edge[:, :, -1] = 1
    u_bc = torch.randn(1, 1, H, W)
    xg, yg = fem.xgp, fem.ygp
    f_gp = (2 * np.pi ** 2 * torch.sin(np.pi * xg) * torch.sin(np.pi * yg)).float()
    kw = dict(nu=nu, f_gp=f_gp, dirichlet=[(edge, u_bc)], c_k=0.5)
    loss, grad = run_energy(fem, u, **kw)
    lref, gref = oracle_energy(fem, u, **kw)
    assert_parity(loss, grad, lref, gref, masks=(edge.expand(B, 1, H, W),), what="vf+fgp")


@pytest.mark.parametrize("ngp", [3, 4])
def test_more_gauss_points(ngp):
    """ngp_1d = 3, 4: the closed form uses the second moment of the reference's own rule."""
    B, H, W = 2, 20, 28
    fem = DiffNet2DFEM(None, domain_sizes=(W, H, 1), domain_size=W, ngp_1d=ngp)
    u, inputs, f = make_inputs(B, H, W, seed=ngp)
    kw = dict(nu=inputs[:, 0:1], f=f, dirichlet=[(inputs[:, 1:2], 1.0), (inputs[:, 2:3], 0.0)])
    loss, grad = run_energy(fem, u, **kw)
    lref, gref = oracle_energy(fem, u, **kw)
    assert_parity(loss, grad, lref, gref, what=f"ngp{ngp}")
    f_gp = torch.randn(1, ngp * ngp, H - 1, W - 1)
    kw = dict(nu=inputs[:, 0:1], f_gp=f_gp, c_k=0.5)
    loss, grad = run_energy(fem, u, **kw)
    lref, gref = oracle_energy(fem, u, **kw)
    assert_parity(loss, grad, lref, gref, what=f"ngp{ngp} fgp")


def test_strided_channel_slices_and_broadcast():
    """nu/bc1/bc2 as channel slices of one (B,3,H,W) tensor (no .contiguous()), f broadcast
    over the batch, u a bare (H,W) parameter."""
    B, H, W = 4, 32, 64
    fem = DiffNet2DFEM(None, domain_sizes=(W, H, 1), domain_size=W)
    u, inputs, f = make_inputs(B, H, W, seed=5)
    inp = inputs.to(DEV)
    f1 = f[:1]
    ud = u.to(DEV).requires_grad_(True)
    loss = fem.energy_loss(ud, nu=inp[:, 0:1], f=f1.to(DEV),
                           dirichlet=[(inp[:, 1:2], 1.0), (inp[:, 2:3], 0.0)])
    loss.backward()
    lref, gref = oracle_energy(fem, u, nu=inputs[:, 0:1], f=f1,
                               dirichlet=[(inputs[:, 1:2], 1.0), (inputs[:, 2:3], 0.0)])
    assert_parity(loss.cpu(), ud.grad.cpu(), lref, gref, what="strided")
    # bare (H, W) parameter against batched inputs: gradient is summed over the batch
    ub = u[0, 0].clone()
    ubd = ub.to(DEV).requires_grad_(True)
    loss = fem.energy_loss(ubd, nu=inp[:, 0:1], f=f1.to(DEV), dirichlet=[(inp[:, 1:2], 1.0)])
    loss.backward()
    lref, gref = oracle_energy(fem, ub, nu=inputs[:, 0:1], f=f1, dirichlet=[(inputs[:, 1:2], 1.0)])
    assert tuple(ubd.grad.shape) == (H, W)
    assert_parity(loss.cpu(), ubd.grad.cpu(), lref, gref, what="bare u")


def test_grad_output_scaling_and_grad_nu():
    """backward multiplies by grad_output; d loss / d nu (16_topopt.py:124,153)."""
    B, H, W = 2, 28, 36
    fem = DiffNet2DFEM(None, domain_sizes=(W, H, 1), domain_size=W)
    u, inputs, f = make_inputs(B, H, W, seed=9)
    nu, bc1, bc2 = inputs[:, 0:1], inputs[:, 1:2], inputs[:, 2:3]
    ud = u.to(DEV).requires_grad_(True)
    nud = nu.to(DEV).clone().requires_grad_(True)
    loss = 3.5 * fem.energy_loss(ud, nu=nud, f=f.to(DEV), dirichlet=[(bc1.to(DEV), 1.0), (bc2.to(DEV), 0.0)])
    loss.backward()
    from helpers import oracle_for, to64
    from oracle import losses as OL
    o = oracle_for(fem)
    u64, nu64 = to64(u).requires_grad_(True), to64(nu).requires_grad_(True)
    lref = 3.5 * OL.energy_loss(o, u64, nu=nu64, f=to64(f), dirichlet=[(to64(bc1), 1.0), (to64(bc2), 0.0)])
    gu, gn = torch.autograd.grad(lref, (u64, nu64))
    assert rel_scalar(loss.cpu(), lref) <= LOSS_RTOL
    assert rel_l2(ud.grad.cpu(), gu) <= GRAD_RTOL
    assert rel_l2(nud.grad.cpu(), gn) <= GRAD_RTOL


def test_golden_vectors(golden):
    """Outputs of the real reference (tests/golden/make_golden.py)."""
    g = golden("ref_2d_rect")
    X, Y = (int(v) for v in g["sizes"])
    fem = DiffNet2DFEM(None, domain_sizes=(X, Y, 1), domain_lengths=(1.5, 1.0, 1.0), domain_size=X,
                       domain_length=1.5)
    u, inputs, f = g.t("u"), g.t("inputs"), g.t("forcing")
    nu, bc1, bc2 = inputs[:, 0:1], inputs[:, 1:2], inputs[:, 2:3]
    d = [(bc1, 1.0), (bc2, 0.0)]
    for key, kw in (("E1", dict(nu=nu, f=f, dirichlet=d)),
                    ("E2", dict(nu=nu, f=f, dirichlet=d, scale=0.5 * (0.5 * fem.h) ** 2)),
                    ("E3fgp", dict(nu=nu, f_gp=g.t("f_gp"), dirichlet=[(bc2, g.t("u_bc"))], c_k=0.5))):
        loss, grad = run_energy(fem, u, **kw)
        assert rel_scalar(loss, g[key + ".loss64"]) <= LOSS_RTOL, key
        assert rel_l2(grad, g[key + ".grad64"]) <= GRAD_RTOL, key
        assert rel_scalar(loss, g[key + ".loss"]) <= LOSS_RTOL, key
    g = golden("ref_2d_neumann")
    fem = DiffNet2DFEM(None, domain_size=17)
    u, inputs, f = g.t("u"), g.t("inputs"), g.t("forcing")
    nu, obj, bc2, bc3 = (inputs[:, i:i + 1] for i in range(4))
    loss, grad = run_energy(fem, u, nu=nu, f=f, nu_zero_mask=obj, dirichlet=[(bc2, 1.0), (bc3, 0.0)])
    assert rel_scalar(loss, g["E5.loss64"]) <= LOSS_RTOL and rel_l2(grad, g["E5.grad64"]) <= GRAD_RTOL
    g = golden("ref_2d_ngp")
    u, inputs, f = g.t("u"), g.t("inputs"), g.t("forcing")
    for ngp in (3, 4):
        fem = DiffNet2DFEM(None, domain_size=10, ngp_1d=ngp)
        loss, grad = run_energy(fem, u, nu=inputs[:, 0:1], f=f,
                                dirichlet=[(inputs[:, 1:2], 1.0), (inputs[:, 2:3], 0.0)])
        assert rel_scalar(loss, g[f"E1_ngp{ngp}.loss64"]) <= LOSS_RTOL
        assert rel_l2(grad, g[f"E1_ngp{ngp}.grad64"]) <= GRAD_RTOL


def test_residual_form(golden):
    """Assembled residual sum(R^2) and its gradient (12_klsum.py:80-132, tests/test.py)."""
    B, H, W = 2, 26, 40
    fem = DiffNet2DFEM(None, domain_sizes=(W, H, 1), domain_size=W)
    u, inputs, f = make_inputs(B, H, W, seed=13, smooth=True)
    nu, bc1, bc2 = inputs[:, 0:1], inputs[:, 1:2], inputs[:, 2:3]
    d = [(bc1, 1.0), (bc2, 0.0)]
    ud = u.to(DEV).requires_grad_(True)
    loss = fem.residual_loss(ud, nu=nu.to(DEV), f=f.to(DEV), dirichlet=[(m.to(DEV), v) for m, v in d],
                             jac=1.0)
    loss.backward()
    lref, gref = oracle_residual(fem, u, nu=nu, f=f, dirichlet=d, jac=1.0)
    assert_parity(loss.cpu(), ud.grad.cpu(), lref, gref, masks=(bc1, bc2), what="resmin")
    g = golden("ref_2d_rect")
    X, Y = (int(v) for v in g["sizes"])
    fem = DiffNet2DFEM(None, domain_sizes=(X, Y, 1), domain_lengths=(1.5, 1.0, 1.0), domain_size=X,
                       domain_length=1.5)
    u, inputs, f = g.t("u"), g.t("inputs"), g.t("forcing")
    ud = u.to(DEV).requires_grad_(True)
    inp = inputs.to(DEV)
    loss = fem.residual_loss(ud, nu=inp[:, 0:1], f=f.to(DEV), dirichlet=[(inp[:, 1:2], 1.0), (inp[:, 2:3], 0.0)])
    loss.backward()
    assert rel_scalar(loss.cpu(), g["resmin.loss64"]) <= LOSS_RTOL
    assert rel_l2(ud.grad.cpu(), g["resmin.grad64"]) <= GRAD_RTOL


def test_gauss_point_evaluation_and_adjoint(golden):
    """gauss_pt_evaluation{,_der_x,_der_y} + their backward vs the reference's conv outputs."""
    g = golden("ref_2d_rect")
    X, Y = (int(v) for v in g["sizes"])
    fem = DiffNet2DFEM(None, domain_sizes=(X, Y, 1), domain_lengths=(1.5, 1.0, 1.0), domain_size=X,
                       domain_length=1.5)
    u = g.t("u").to(DEV)
    for key, fn in (("N", fem.gauss_pt_evaluation), ("dx", fem.gauss_pt_evaluation_der_x),
                    ("dy", fem.gauss_pt_evaluation_der_y)):
        out = fn(u)
        assert out.shape == g["gp." + key].shape
        assert rel_l2(out.cpu(), g["gp." + key]) <= 1e-6, key
    # a user-written (un-fused) loss body on the CUDA gp-eval ops == oracle autograd
    B, H, W = 2, 19, 23
    fem = DiffNet2DFEM(None, domain_sizes=(W, H, 1), domain_size=W, ngp_1d=3)
    u, inputs, f = make_inputs(B, H, W, seed=21)
    ud = u.to(DEV).requires_grad_(True)
    nu_gp = fem.gauss_pt_evaluation(inputs[:, 0:1].to(DEV))
    ux, uy, ug = fem.gauss_pt_evaluation_der_x(ud), fem.gauss_pt_evaluation_der_y(ud), fem.gauss_pt_evaluation(ud)
    w = fem.gpw.to(DEV)[None, :, None, None]
    loss = torch.mean(torch.sum(w * (nu_gp * (ux ** 2 + uy ** 2) - ug * fem.gauss_pt_evaluation(f.to(DEV))), 1))
    loss.backward()
    lref, gref = oracle_energy(fem, u, nu=inputs[:, 0:1], f=f)
    assert_parity(loss.cpu(), ud.grad.cpu(), lref, gref, what="unfused body")


def test_full_size_properties():
    """BASELINE sizes (256^2 B=64, 512^2 B=16): size-independent properties instead of the
    oracle -- Euler identity <grad,u> = 2E for the pure stiffness energy, exact zero energy and
    gradient for constant u, quadratic scaling, chunking independence, and run-to-run bit
    reproducibility."""
    for B, N in ((64, 256), (16, 512)):
        fem = DiffNet2DFEM(None, domain_size=N, batch_size=B)
        g = torch.Generator(device=DEV).manual_seed(N)
        u = torch.randn(B, 1, N, N, device=DEV, generator=g)
        nu = torch.exp(0.3 * torch.randn(B, 1, N, N, device=DEV, generator=g))
        loss, grad = fem.energy_loss_and_grad(u, nu=nu)
        euler = float((grad.double() * u[:, 0].double()).sum())
        assert rel_scalar(euler, 2.0 * float(loss)) < 2e-5
        l2, g2 = fem.energy_loss_and_grad(2.0 * u, nu=nu)
        assert rel_scalar(l2, 4.0 * float(loss)) < 1e-5 and rel_l2(g2, 2.0 * grad) < 1e-5
        lc, gc = fem.energy_loss_and_grad(torch.full_like(u, 3.0), nu=nu)
        assert float(lc) == 0.0 and float(gc.abs().max()) == 0.0
        la, ga = fem.energy_loss_and_grad(u, nu=nu)
        assert torch.equal(la, loss) and torch.equal(ga, grad)            # deterministic
        saved = {k: os.environ.get(k) for k in ("DN_R_2D", "DN_T2_R")}
        try:
            os.environ["DN_R_2D"] = "37"                                   # different row chunking
            os.environ["DN_T2_R"] = "37"
            lb, gb = fem.energy_loss_and_grad(u, nu=nu)
        finally:
            for k, v in saved.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
        assert rel_scalar(lb, loss) < 1e-6 and torch.equal(gb, grad)       # seams are exact
        # batch samples are independent: sample 3 alone gives its slice of the gradient
        l1, g1 = fem.energy_loss_and_grad(u[3:4], nu=nu[3:4], reduction="sum")
        ls, gs = fem.energy_loss_and_grad(u, nu=nu, reduction="sum")
        assert torch.equal(g1[0], gs[3])


@pytest.mark.parametrize("B,N", [(2, 256), (1, 512)])
def test_baseline_grids_against_oracle(B, N):
    """The BASELINE grids themselves (256^2, 512^2; small batch so the fp64 oracle finishes in
    seconds) through the streaming kernel: loss <= 1e-5, gradient <= 1e-4, zeros on Dirichlet nodes."""
    fem = DiffNet2DFEM(None, domain_size=N)
    u, inputs, f = make_inputs(B, N, N, seed=N)
    nu, bc1, bc2 = inputs[:, 0:1], inputs[:, 1:2], inputs[:, 2:3]
    kw = dict(nu=nu, f=f, dirichlet=[(bc1, 1.0), (bc2, 0.0)])
    loss, grad = run_energy(fem, u, **kw)
    lref, gref = oracle_energy(fem, u, **kw)
    assert_parity(loss, grad, lref, gref, masks=(bc1, bc2), what=f"E1 {B}x{N}x{N}")


@pytest.mark.parametrize("B,N,samples", [(64, 256, (0, 31, 63)), (16, 512, (0, 7, 15))])
def test_bench_launch_shapes_against_oracle(B, N, samples):
    """The launch shapes the bench times (256^2 x 64: one wave of short row chunks; 512^2 x 16), not a
    small-batch stand-in: the launch plan depends on B, samples are independent, so three samples of
    the full-batch launch are compared with the fp64 oracle run on those samples alone."""
    from diffnet_b200.synthetic import poisson2d_parametric_batch
    fem = DiffNet2DFEM(None, domain_size=N, batch_size=B)
    u, inputs, f = poisson2d_parametric_batch(B, N, DEV, seed=B + N)
    kw = dict(nu=inputs[:, 0:1], f=f, dirichlet=[(inputs[:, 1:2], 1.0), (inputs[:, 2:3], 0.0)])
    loss, grad = fem.energy_loss_and_grad(u, reduction="sum", **kw)
    tot = 0.0
    for b in samples:
        sl = slice(b, b + 1)
        kwb = dict(nu=inputs[sl, 0:1], f=f[sl], dirichlet=[(inputs[sl, 1:2], 1.0), (inputs[sl, 2:3], 0.0)])
        lref, gref = oracle_energy(fem, u[sl], reduction="sum", **kwb)
        lb, gb = fem.energy_loss_and_grad(u[sl], reduction="sum", **kwb)      # B = 1 launch of the same sample
        assert_parity(lb, grad[b], lref, gref[0, 0], masks=(inputs[b, 1], inputs[b, 2]),
                      what=f"sample {b} of the B={B} launch at {N}^2")
        assert torch.equal(gb[0], grad[b])                                    # seams are exact: plans agree bit for bit
        tot += float(lref)
    assert math.isfinite(float(loss))


def test_immersed_geometry_masks():
    """configs[4]: IBN 2-D with irregular (star-shaped) object masks, nu = domain indicator
    (0 inside the object), bc1 = object, bc2 = the four edges -- overlapping masks on the corners and
    wherever a silhouette touches an edge (e1_complex_immersed_background.py:33-58)."""
    from diffnet_b200.synthetic import ibn2d_batch
    B, N = 3, 128
    fem = DiffNet2DFEM(None, domain_size=N)
    u, inputs, f = ibn2d_batch(B, N, torch.device("cpu"), seed=3)
    f = torch.randn_like(f)
    nu, bc1, bc2 = inputs[:, 0:1], inputs[:, 1:2], inputs[:, 2:3]
    assert 0.02 < float(bc1.mean()) < 0.6
    kw = dict(nu=nu, f=f, dirichlet=[(bc1, 1.0), (bc2, 0.0)])
    loss, grad = run_energy(fem, u, **kw)
    lref, gref = oracle_energy(fem, u, **kw)
    assert_parity(loss, grad, lref, gref, masks=(bc1, bc2), what="IBN 2-D silhouettes")


def test_streaming_and_warp_paths_agree():
    """The bulk-async streaming kernel (k_fem2d_tma) and the general warp-marching kernel
    (k_fem2d) are two implementations of the same operator: same loss, same gradient, on the
    aligned sizes both accept, for every Dirichlet-set / nu / f combination."""
    B, H, W = 3, 70, 264
    fem = DiffNet2DFEM(None, domain_sizes=(W, H, 1), domain_size=W)
    u, inputs, f = make_inputs(B, H, W, seed=5)
    u, inputs, f = u.to(DEV), inputs.to(DEV), f.to(DEV)
    nu, bc1, bc2 = inputs[:, 0:1], inputs[:, 1:2], inputs[:, 2:3]
    bc3 = torch.zeros_like(bc1); bc3[:, :, 0, :] = 1
    ubc = torch.randn_like(u)
    cases = [dict(), dict(nu=nu), dict(f=f), dict(nu=nu, f=f, dirichlet=[(bc1, 1.0)]),
             dict(nu=nu, f=f, dirichlet=[(bc1, 1.0), (bc2, 0.0)]),
             dict(f=f, dirichlet=[(bc1, 1.0), (bc2, 0.0), (bc3, 0.25)]),
             dict(nu=nu, f=f, dirichlet=[(bc2, ubc)]),
             dict(nu=nu, f=f, nu_zero_mask=bc3, dirichlet=[(bc1, 1.0), (bc2, 0.0)])]
    for kw in cases:
        os.environ.pop("DN_2D_PATH", None)
        ls, gs = fem.energy_loss_and_grad(u, **kw)
        os.environ["DN_2D_PATH"] = "warp"
        try:
            lw, gw = fem.energy_loss_and_grad(u, **kw)
        finally:
            os.environ.pop("DN_2D_PATH", None)
        assert rel_scalar(ls, lw) < 2e-6, (sorted(kw), float(ls), float(lw))
        assert rel_l2(gs, gw) < 2e-6, sorted(kw)
        assert torch.equal(gs == 0, gw == 0) or rel_l2(gs, gw) < 2e-6


def test_residual_form_streaming_vs_general():
    """sum(R^2) and its gradient (one more operator pass with mask_input = 0) on the streaming
    kernels vs the general kernels."""
    B, H, W = 2, 40, 136
    fem = DiffNet2DFEM(None, domain_sizes=(W, H, 1), domain_size=W)
    u, inputs, f = make_inputs(B, H, W, seed=9)
    u, inputs, f = u.to(DEV), inputs.to(DEV), f.to(DEV)
    nu, bc1, bc2 = inputs[:, 0:1], inputs[:, 1:2], inputs[:, 2:3]
    out = {}
    for path in ("", "warp"):
        if path:
            os.environ["DN_2D_PATH"] = path
        try:
            ud = u.clone().requires_grad_(True)
            loss = fem.residual_loss(ud, nu=nu, f=f, dirichlet=[(bc1, 1.0), (bc2, 0.0)], jac=(0.5 * fem.h) ** 2)
            loss.backward()
            out[path] = (loss.detach(), ud.grad.detach())
        finally:
            os.environ.pop("DN_2D_PATH", None)
    assert rel_scalar(out[""][0], out["warp"][0]) < 2e-6
    assert rel_l2(out[""][1], out["warp"][1]) < 2e-6


def test_prepared_call_equals_plain_call():
    """ops.PreparedEnergy (marshalling done once) == energy_loss_and_grad, sees in-place updates of
    the bound tensors, and reuses its output buffers."""
    fem = DiffNet2DFEM(None, domain_size=64)
    u, inputs, f = make_inputs(3, 64, 64, seed=2)
    u, inputs, f = u.to(DEV), inputs.to(DEV), f.to(DEV)
    kw = dict(nu=inputs[:, 0:1], f=f, dirichlet=[(inputs[:, 1:2], 1.0), (inputs[:, 2:3], 0.0)], c_k=0.5)
    call = fem.prepare_energy(u, **kw)
    l0, g0 = fem.energy_loss_and_grad(u, **kw)
    l1, g1 = call()
    assert torch.equal(l0, l1) and torch.equal(g0, g1)
    u.mul_(1.5)
    l2, g2 = fem.energy_loss_and_grad(u, **kw)
    l3, g3 = call()
    assert l3.data_ptr() == l1.data_ptr() and g3.data_ptr() == g1.data_ptr()
    assert torch.equal(l2, l3) and torch.equal(g2, g3)


def test_concurrent_streams_and_graph_replay():
    """Two CUDA streams issue launches concurrently (one workspace per stream), and a captured
    graph of several launches replays to the same bits as eager execution."""
    fem = DiffNet2DFEM(None, domain_size=128)
    u, inputs, f = make_inputs(8, 128, 128, seed=4)
    u, inputs, f = u.to(DEV), inputs.to(DEV), f.to(DEV)
    kw = dict(nu=inputs[:, 0:1], f=f, dirichlet=[(inputs[:, 1:2], 1.0), (inputs[:, 2:3], 0.0)])
    lref, gref = fem.energy_loss_and_grad(u, **kw)
    u2 = 2.0 * u
    lref2, gref2 = fem.energy_loss_and_grad(u2, **kw)
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    outs1, outs2 = [], []
    for _ in range(20):
        with torch.cuda.stream(s1):
            outs1.append(fem.energy_loss_and_grad(u, **kw))
        with torch.cuda.stream(s2):
            outs2.append(fem.energy_loss_and_grad(u2, **kw))
    torch.cuda.synchronize()
    for (l, g), (l2, g2) in zip(outs1, outs2):
        assert torch.equal(l, lref) and torch.equal(g, gref)
        assert torch.equal(l2, lref2) and torch.equal(g2, gref2)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fem.energy_loss_and_grad(u, **kw)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        captured = [fem.energy_loss_and_grad(u if i % 2 == 0 else u2, **kw) for i in range(6)]
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    for i, (l, g) in enumerate(captured):
        assert torch.equal(l, lref if i % 2 == 0 else lref2) and torch.equal(g, gref if i % 2 == 0 else gref2)


def test_errors_are_loud():
    from diffnet_b200._lib import DiffNetFEMError
    fem = DiffNet2DFEM(None, domain_size=16)
    u = torch.zeros(1, 1, 16, 16, device=DEV)
    with pytest.raises(DiffNetFEMError):
        fem.energy_loss(torch.zeros(1, 1, 15, 16, device=DEV))
    with pytest.raises(DiffNetFEMError):
        fem.energy_loss(u.double())
    with pytest.raises(DiffNetFEMError):
        fem.energy_loss(u, f=u, f_gp=torch.zeros(1, 4, 15, 15, device=DEV))
    with pytest.raises(DiffNetFEMError):
        fem.energy_loss(u, dirichlet=[(u, 0.0)] * 4)
    with pytest.raises(DiffNetFEMError):
        fem.energy_loss(u.cpu())


@pytest.mark.parametrize("B,H,W,ngp", [(3, 70, 300, 2), (2, 33, 64, 2), (1, 9, 40, 3), (2, 17, 31, 4)])
def test_gp_eval_marching_kernels_against_oracle(B, H, W, ngp):
    """The marching gp-eval kernels (several x chunks / y chunks / warp seams of the adjoint) and the
    multi-table pass: forward values and the adjoint (through autograd) against the oracle's convs in fp64."""
    from helpers import oracle_for
    fem = DiffNet2DFEM(None, domain_sizes=(W, H, 1), domain_lengths=(1.3, 0.9, 1.0), domain_size=W, ngp_1d=ngp)
    o = oracle_for(fem)
    g = torch.Generator().manual_seed(B * 1000 + W)
    u = torch.randn(B, 1, H, W, generator=g)
    ud = u.to(DEV).requires_grad_(True)
    outs = fem.gauss_pt_evaluation_all(ud)
    uo = u.double().requires_grad_(True)
    refs = (o.gauss_pt_evaluation(uo), o.gauss_pt_evaluation_der_x(uo), o.gauss_pt_evaluation_der_y(uo))
    single = (fem.gauss_pt_evaluation(ud), fem.gauss_pt_evaluation_der_x(ud), fem.gauss_pt_evaluation_der_y(ud))
    cot = [torch.randn(r.shape, generator=g) for r in refs]
    for a, s1, r in zip(outs, single, refs):
        assert a.shape == r.shape
        assert rel_l2(a.detach().cpu(), r.detach()) <= 1e-6
        assert torch.equal(a, s1)                       # one pass == one launch per table, bit for bit
    sum((a * c.to(DEV)).sum() for a, c in zip(outs, cot)).backward()
    sum((r * c.double()).sum() for r, c in zip(refs, cot)).backward()
    assert rel_l2(ud.grad.cpu(), uo.grad) <= 1e-6
    assert float((ud.grad.cpu().double() - uo.grad).abs().max() / uo.grad.abs().max()) <= 2e-6
    # the single-table adjoints add up to the multi-table one
    u2 = u.to(DEV).requires_grad_(True)
    sum((fn(u2) * c.to(DEV)).sum() for fn, c in zip((fem.gauss_pt_evaluation, fem.gauss_pt_evaluation_der_x,
                                                     fem.gauss_pt_evaluation_der_y), cot)).backward()
    assert rel_l2(u2.grad, ud.grad) <= 1e-6


@pytest.mark.parametrize("B,H,W", [(40, 128, 64), (64, 256, 64), (100, 300, 32), (23, 510, 128), (16, 512, 256)])
def test_balanced_one_wave_split_against_the_general_kernel(B, H, W):
    """Launch shapes the one-wave planner cuts into equal parts per image (a subset of the slots, interior cuts
    at odd rows, images with q and q + 1 chunks): every row must be owned by exactly one chunk.  Cross-checked
    against the general kernel (a different decomposition) and against the forced uniform split."""
    fem = DiffNet2DFEM(None, domain_sizes=(W, H, 1), domain_size=W)
    g = torch.Generator().manual_seed(B * 7 + H)
    u = torch.randn(B, 1, H, W, generator=g).to(DEV)
    nu = torch.exp(0.3 * torch.randn(B, 1, H, W, generator=g)).to(DEV)
    f = torch.randn(B, 1, H, W, generator=g).to(DEV)
    bc = (torch.rand(B, 1, H, W, generator=g) > 0.9).float().to(DEV)
    kw = dict(nu=nu, f=f, dirichlet=[(bc, 0.5)])
    ls, gs = fem.energy_loss_and_grad(u, **kw)
    for knobs in ({"DN_2D_PATH": "warp"}, {"DN_T2_BALANCE": "0"}, {"DN_T2_FILL_PCT": "100"}, {"DN_T2_FILL_PCT": "45"}):
        os.environ.update(knobs)
        try:
            lo, go = fem.energy_loss_and_grad(u, **kw)
        finally:
            for k in knobs:
                os.environ.pop(k)
        assert rel_scalar(ls, lo) < 2e-6, knobs
        assert rel_l2(gs, go) < 2e-6 and float((gs - go).abs().max() / go.abs().max()) < 1e-5, knobs
