"""The oracle vs. the committed golden vectors (outputs of the real reference made by
tests/golden/make_golden.py).  This is the pinning that travels to the GPU box."""
import numpy as np
import pytest
import torch

from conftest import rel_l2, rel_scalar
from oracle import losses as L
from oracle.fem import Q1Oracle

torch.set_num_threads(max(1, min(4, torch.get_num_threads())))


def _lg(fn, u):
    u = u.clone().requires_grad_(True)
    loss = fn(u)
    (g,) = torch.autograd.grad(loss, u)
    return loss.detach(), g


def _check(g, key, loss, grad, exact=True):
    if exact:
        # bit-identical on the CPU that generated the vectors; another CPU's oneDNN kernels may
        # associate the 4/8-tap sums differently, so fall back to a few-ulp bound there.
        if not (np.array_equal(loss.numpy(), g[key + ".loss"])
                and np.array_equal(grad.numpy(), g[key + ".grad"])):
            assert rel_scalar(loss, g[key + ".loss"]) < 2e-6, key
            assert rel_l2(grad, g[key + ".grad"]) < 2e-6, key
    assert rel_scalar(loss, g[key + ".loss64"]) < 1e-5
    assert rel_l2(grad, g[key + ".grad64"]) < 1e-4


def test_2d_rect(golden):
    g = golden("ref_2d_rect")
    X, Y = (int(v) for v in g["sizes"])
    fem = Q1Oracle(nsd=2, domain_sizes=(X, Y, 1), domain_lengths=(1.5, 1.0, 1.0),
                   domain_size=X, domain_length=1.5)
    u, inputs, f = g.t("u"), g.t("inputs"), g.t("forcing")
    assert fem.h == float(g["h"]) and fem.hs == (float(g["hx"]), float(g["hy"]))
    assert np.array_equal(fem.gauss_pt_evaluation(u).numpy(), g["gp.N"])
    assert np.array_equal(fem.gauss_pt_evaluation_der_x(u).numpy(), g["gp.dx"])
    assert np.array_equal(fem.gauss_pt_evaluation_der_y(u).numpy(), g["gp.dy"])
    _check(g, "E1", *_lg(lambda v: L.body_klsum_energy(fem, v, inputs, f), u))
    _check(g, "E2", *_lg(lambda v: L.body_0_base(fem, v, inputs, f), u))
    _check(g, "resmin", *_lg(lambda v: L.body_klsum_resmin(fem, v, inputs, f), u))
    nu, bc2 = inputs[:, 0:1], inputs[:, 2:3]
    _check(g, "E3fgp", *_lg(lambda v: L.energy_loss(
        fem, v, nu=nu, f_gp=g.t("f_gp"), dirichlet=[(bc2, g.t("u_bc"))], c_k=0.5), u))
    # the parametrised form equals the literal bodies (same ops up to association)
    bc1 = inputs[:, 1:2]
    l, gr = _lg(lambda v: L.energy_loss(fem, v, nu=nu, f=f, dirichlet=[(bc1, 1.0), (bc2, 0.0)]), u)
    _check(g, "E1", l, gr, exact=False)
    l, gr = _lg(lambda v: L.energy_loss(fem, v, nu=nu, f=f, dirichlet=[(bc1, 1.0), (bc2, 0.0)],
                                        scale=0.5 * (0.5 * fem.h) ** 2), u)
    _check(g, "E2", l, gr, exact=False)
    l, gr = _lg(lambda v: L.residual_loss(fem, v, nu=nu, f=f, dirichlet=[(bc1, 1.0), (bc2, 0.0)]), u)
    _check(g, "resmin", l, gr, exact=False)


def test_2d_neumann(golden):
    g = golden("ref_2d_neumann")
    fem = Q1Oracle(nsd=2, domain_size=17)
    u, inputs, f = g.t("u"), g.t("inputs"), g.t("forcing")
    _check(g, "E5", *_lg(lambda v: L.body_ibn2d_neumann(fem, v, inputs, f), u))
    nu, obj, bc2, bc3 = (inputs[:, i:i + 1] for i in range(4))
    l, gr = _lg(lambda v: L.energy_loss(fem, v, nu=nu, f=f, nu_zero_mask=obj,
                                        dirichlet=[(bc2, 1.0), (bc3, 0.0)]), u)
    _check(g, "E5", l, gr, exact=False)


@pytest.mark.parametrize("ngp", [3, 4])
def test_2d_more_gauss_points(golden, ngp):
    g = golden("ref_2d_ngp")
    fem = Q1Oracle(nsd=2, domain_size=10, ngp_1d=ngp)
    u, inputs, f = g.t("u"), g.t("inputs"), g.t("forcing")
    assert np.array_equal(fem.gpw.numpy(), g[f"gpw_ngp{ngp}"])
    assert np.array_equal(fem.gauss_pt_evaluation(u).numpy(), g[f"gp.N_ngp{ngp}"])
    assert np.array_equal(fem.gauss_pt_evaluation_der_x(u).numpy(), g[f"gp.dx_ngp{ngp}"])
    _check(g, f"E1_ngp{ngp}", *_lg(lambda v: L.body_klsum_energy(fem, v, inputs, f), u))


def test_3d_box(golden):
    g = golden("ref_3d_box")
    X, Y, Z = (int(v) for v in g["sizes"])
    fem = Q1Oracle(nsd=3, domain_sizes=(X, Y, Z), domain_lengths=(1.0, 0.8, 0.5), domain_size=X)
    u, inputs, f = g.t("u"), g.t("inputs"), g.t("forcing")
    for k, m in (("N", "gauss_pt_evaluation"), ("dx", "gauss_pt_evaluation_der_x"),
                 ("dy", "gauss_pt_evaluation_der_y"), ("dz", "gauss_pt_evaluation_der_z")):
        assert np.array_equal(getattr(fem, m)(u).numpy(), g["gp." + k]), k
    src, sink = g.t("source"), g.t("sink")
    _check(g, "ibn3d", *_lg(lambda v: L.body_ibn3d(fem, v, src, sink, f), u))
    _check(g, "inobj", *_lg(lambda v: L.body_solve_in_object(fem, v, inputs, f), u))
    _check(g, "inobj_bare", *_lg(lambda v: L.body_solve_in_object(fem, v, inputs[:1], f[:1]), u[0, 0]))
    # parametrised equivalents; IBN_3D's mask fix == "source wins over sink"
    l, gr = _lg(lambda v: L.energy_loss(fem, v, f=f, dirichlet=[(sink, 0.0), (src, 1.0)]), u)
    _check(g, "ibn3d", l, gr, exact=False)
    l, gr = _lg(lambda v: L.energy_loss(fem, v, nu=inputs[:, 0:1], f=f,
                                        dirichlet=[(inputs[:, 1:2], 0.0)], c_k=0.5), u)
    _check(g, "inobj", l, gr, exact=False)


def test_reference_test_scripts(golden):
    """tests/test.py / tests/test3D.py of the reference (constructor bug fixed)."""
    g = golden("ref_tests")
    fem2 = Q1Oracle(nsd=2, domain_size=16)
    _check(g, "res2d", *_lg(lambda v: L.body_test2d_residual(fem2, v, g.t("k2")), g.t("u2")))
    fem3 = Q1Oracle(nsd=3, domain_size=8)
    _check(g, "res3d", *_lg(lambda v: L.body_test3d_residual(fem3, v, g.t("k3")), g.t("u3")))


def test_host_kl_recipe_against_the_reference_fields():
    """synthetic.kl_omegas / kl_diffusivity_2d (host side of the KL producer) against the reference's table and
    fields (tests/golden/producers.npz, made from DiffNet/gen_input_calc.py by make_golden_producers.py)."""
    import os
    import numpy as np
    from conftest import ROOT
    from diffnet_b200.synthetic import kl_diffusivity_2d, kl_omegas
    g = np.load(os.path.join(ROOT, "tests", "golden", "producers.npz"))
    assert np.abs(kl_omegas(0.5, 6) - g["kl.omega"][:6]).max() < 1e-12
    nu = kl_diffusivity_2d(torch.from_numpy(g["kl.coeffs"]), 16)
    ref = torch.from_numpy(g["kl2d.inputs"][:, 0:1])
    assert float(((nu - ref).abs() / ref).max()) < 1e-6


@pytest.mark.parametrize("nsd,sizes,ngp", [(2, (9, 7, 1), 2), (2, (6, 8, 1), 3), (3, (5, 4, 6), 2), (3, (4, 5, 4), 4)])
def test_load_vector_identity_on_the_oracle(nsd, sizes, ngp):
    """What the load-vector route (include/diffnet_fem.h: DN_F_LOAD_VECTOR) relies on, checked on the CPU oracle alone:
    the reference's forcing term sum_g w_g f_g u_g (e8_2d_poisson_mms.py:154-175, u_gp from gauss_pt_evaluation) equals
    sum_a b_a u_a with b_a = sum over the elements around node a and their Gauss points of w_g N_a(g) f_g -- the
    closed form the CUDA assembly implements (1-D factors w_q (1 -+ x_q) / 2 per direction)."""
    import numpy as np
    from oracle.fem import Q1Oracle, gauss_rule
    o = Q1Oracle(nsd=nsd, domain_sizes=sizes, domain_lengths=(1.0, 0.8, 0.6), domain_size=sizes[0], ngp_1d=ngp,
                 dtype=torch.float64)
    nx, ny, nz = sizes
    sp = (ny, nx) if nsd == 2 else (nz, ny, nx)
    el = tuple(d - 1 for d in sp)
    g = torch.Generator().manual_seed(nsd * 10 + ngp)
    u = torch.randn((2, 1) + sp, generator=g, dtype=torch.float64, requires_grad=True)
    f_gp = torch.randn((2, ngp ** nsd) + el, generator=g, dtype=torch.float64)
    w = o.gpw.double().reshape((1, -1) + (1,) * nsd)
    term = (w * f_gp * o.gauss_pt_evaluation(u)).sum()
    (b_ref,) = torch.autograd.grad(term, u)
    gx, gw = gauss_rule(ngp)
    gx, gw = np.asarray(gx, dtype=np.float64), np.asarray(gw, dtype=np.float64)
    wn = np.stack([gw * 0.5 * (1 - gx), gw * 0.5 * (1 + gx)])            # [side][q]
    F = f_gp.numpy().reshape((2,) + (ngp,) * nsd + el)                    # [b][kg][jg][ig][(ez,) ey, ex]
    b = np.zeros((2,) + sp)
    sides = [(0, slice(0, -1)), (1, slice(1, None))]                       # local node 0 -> element index = node index
    if nsd == 2:
        for jb, ys in sides:
            for ib, xs in sides:
                b[:, ys, xs] += np.einsum("j,i,bjiyx->byx", wn[jb], wn[ib], F)
    else:
        for kb, zs in sides:
            for jb, ys in sides:
                for ib, xs in sides:
                    b[:, zs, ys, xs] += np.einsum("k,j,i,bkjizyx->bzyx", wn[kb], wn[jb], wn[ib], F)
    err = np.abs(b - b_ref[:, 0].numpy()).max() / np.abs(b).max()
    assert err < 1e-6, err          # the oracle's stencils are the reference's fp32 tables: equal to fp32 rounding
    t = float(term.detach())
    assert abs(t - float((torch.from_numpy(b) * u.detach()[:, 0]).sum())) <= 1e-6 * abs(t) + 1e-9


def test_3d_forcing_at_gauss_points(golden):
    """examples/poisson/mms/e8_3d_poisson_mms.py form (Dirichlet value field + f at the Gauss points) on the real
    reference's 3-D module (tests/golden/make_golden_fgp3d.py): the oracle's energy_loss with f_gp reproduces it."""
    g = golden("ref_3d_fgp")
    X, Y, Z = (int(v) for v in g["sizes"])
    fem = Q1Oracle(nsd=3, domain_sizes=(X, Y, Z), domain_lengths=tuple(float(v) for v in g["lengths"]), domain_size=X)
    _check(g, "E3fgp", *_lg(lambda v: L.energy_loss(fem, v, nu=g.t("nu"), f_gp=g.t("f_gp"),
                                                     dirichlet=[(g.t("bc"), g.t("u_bc"))], c_k=0.5), g.t("u")))
