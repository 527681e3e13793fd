"""The forcing-at-Gauss-points form through the assembled load vector (include/diffnet_fem.h: DN_F_LOAD_VECTOR).

The reference integrates -f_gp * u_gp with the quadrature weights (e8_2d_poisson_mms.py:154-175).  That term is
linear in u, so it is sum_a b_a u_a with b assembled once from f_gp; the streaming kernels then read b (4 bytes per
node) instead of f_gp (4 ngp bytes per element).  Checked here against the oracle (fp64, the reference's own
f_gp * u_gp form) and against the general kernels that read f_gp directly; tolerances of helpers.py.
"""
import pytest
import torch

from helpers import assert_parity, oracle_energy, oracle_for
from diffnet_b200 import DiffNet2DFEM, DiffNet3DFEM, ops
from diffnet_b200 import _lib as L

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _fem(nsd, sizes, ngp=2):
    if nsd == 2:
        W, H = sizes
        return DiffNet2DFEM(None, domain_sizes=(W, H, 1), domain_lengths=(1.0, 0.7, 1.0), domain_size=W, ngp_1d=ngp)
    W, H, D = sizes
    return DiffNet3DFEM(None, domain_sizes=(W, H, D), domain_lengths=(1.0, 0.7, 0.9), domain_size=W, ngp_1d=ngp)


def _load_vector_ref(fem, f_gp):
    """b = d/du sum_g w_g f_g u_g with the oracle's own Gauss-point evaluation (fp64)."""
    o = oracle_for(fem)
    sp = (fem.geometry.ny, fem.geometry.nx) if fem.nsd == 2 else (fem.geometry.nz, fem.geometry.ny, fem.geometry.nx)
    u = torch.zeros((f_gp.shape[0], 1) + sp, dtype=torch.float64, requires_grad=True)
    w = o.gpw.double().reshape((1, -1) + (1,) * fem.nsd)
    term = (w * f_gp.double() * o.gauss_pt_evaluation(u)).sum()
    (b,) = torch.autograd.grad(term, u)
    return b[:, 0]


@pytest.mark.parametrize("nsd,sizes,ngp,Bf", [(2, (12, 9), 2, 1), (2, (37, 21), 3, 3), (2, (64, 32), 4, 2),
                                              (3, (8, 7, 6), 2, 1), (3, (13, 9, 5), 3, 2)])
def test_load_vector_assembly(nsd, sizes, ngp, Bf):
    fem = _fem(nsd, sizes, ngp)
    g = torch.Generator().manual_seed(nsd * 100 + ngp)
    f_gp = torch.randn((Bf, ngp ** nsd) + fem.geometry.elems, generator=g)
    b = ops.load_vector(fem.geometry, f_gp.to(DEV)).cpu()
    ref = _load_vector_ref(fem, f_gp)
    assert b.shape == ref.shape
    err = float((b.double() - ref).abs().max() / ref.abs().max())
    assert err <= 2e-6, err


def _run(fem, u, use_lv, **kw):
    old = ops.USE_LOAD_VECTOR
    ops.USE_LOAD_VECTOR = use_lv
    try:
        ud = u.to(DEV).requires_grad_(True)
        kwd = {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in kw.items() if k != "dirichlet"}
        kwd["dirichlet"] = [(m.to(DEV), (v.to(DEV) if torch.is_tensor(v) else v)) for m, v in kw.get("dirichlet", [])]
        loss = fem.energy_loss(ud, **kwd)
        loss.backward()
        return loss.detach().cpu(), ud.grad.detach().cpu()
    finally:
        ops.USE_LOAD_VECTOR = old


def _case(nsd, sizes, B, Bf, ngp, with_nu, nmask, seed, value_field=False):
    fem = _fem(nsd, sizes, ngp)
    sp = fem.geometry.spatial
    g = torch.Generator().manual_seed(seed)
    u = torch.randn((B, 1) + sp, generator=g)
    kw = dict(c_k=0.5, f_gp=torch.randn((Bf, ngp ** nsd) + fem.geometry.elems, generator=g))
    if with_nu:
        kw["nu"] = torch.exp(0.5 * torch.randn((B, 1) + sp, generator=g))
    masks = []
    if nmask >= 1:
        m = torch.zeros((B, 1) + sp); m[..., 0] = 1
        masks.append((m, torch.randn((1, 1) + sp, generator=g) if value_field else 1.0))
    if nmask >= 2:
        m = torch.zeros((B, 1) + sp); m[..., -1] = 1; m[:, :, 2:5, 3:6] = 1
        masks.append((m, -0.5))
    kw["dirichlet"] = masks
    return fem, u, kw


CASES = [
    # nsd, sizes,        B, Bf, ngp, nu,    masks, value field
    (2, (64, 48),        3, 1,  2,   True,  2, False),
    (2, (256, 40),       2, 2,  2,   False, 1, False),
    (2, (132, 33),       2, 1,  3,   True,  0, False),
    (2, (44, 30),        2, 1,  2,   True,  1, True),      # e8_2d_poisson_mms.py: where(bc, u_bc, u) + f at the Gauss points
    (2, (516, 17),       1, 1,  4,   False, 2, False),
    (3, (32, 20, 12),    2, 1,  2,   True,  2, False),     # anisotropic spacing: NUK 1
    (3, (16, 16, 16),    2, 2,  2,   False, 1, False),
    (3, (24, 10, 9),     1, 1,  3,   True,  1, True),
    (3, (68, 12, 7),     2, 1,  2,   False, 0, False),
]


@pytest.mark.parametrize("case", CASES, ids=[f"{c[0]}d-{'x'.join(map(str, c[1]))}-b{c[2]}" for c in CASES])
def test_f_gp_through_the_load_vector_against_oracle_and_general_kernels(case):
    nsd, sizes, B, Bf, ngp, with_nu, nmask, vf = case
    fem, u, kw = _case(nsd, sizes, B, Bf, ngp, with_nu, nmask, seed=sum(sizes) + ngp, value_field=vf)
    lref, gref = oracle_energy(fem, u, **kw)
    masks = tuple(m for m, _ in kw["dirichlet"])
    l1, g1 = _run(fem, u, True, **kw)
    assert_parity(l1, g1, lref, gref, masks=masks, what="load vector")
    l0, g0 = _run(fem, u, False, **kw)
    assert_parity(l0, g0, lref, gref, masks=masks, what="general f_gp")


def test_isotropic_3d_grid_takes_the_per_node_scaled_variants():
    """hx == hy == hz: the 3-D kernel moves k out of the element (kscale); the load vector must be un-scaled to match."""
    for with_nu in (True, False):
        fem = DiffNet3DFEM(None, domain_sizes=(20, 20, 20), domain_size=20)
        g = torch.Generator().manual_seed(7)
        u = torch.randn(2, 1, 20, 20, 20, generator=g)
        m = torch.zeros(2, 1, 20, 20, 20); m[:, :, 0] = 1
        kw = dict(c_k=0.5, f_gp=torch.randn(1, 8, 19, 19, 19, generator=g), dirichlet=[(m, 0.25)])
        if with_nu:
            kw["nu"] = torch.exp(0.3 * torch.randn(2, 1, 20, 20, 20, generator=g))
        lref, gref = oracle_energy(fem, u, **kw)
        l1, g1 = _run(fem, u, True, **kw)
        assert_parity(l1, g1, lref, gref, masks=(m,), what=f"iso nu={with_nu}")


def test_launches_the_streaming_kernels_cannot_take_fall_back_to_f_gp():
    """Odd nx: DN_F_LOAD_VECTOR is refused with DN_ENOSTREAM by the C ABI; energy_loss then reads f_gp in the general
    kernels -- same numbers."""
    fem, u, kw = _case(2, (37, 21), 2, 1, 2, True, 1, seed=5)
    geom = fem.geometry
    b = ops.load_vector(geom, kw["f_gp"].to(DEV))
    with pytest.raises(L.DiffNetFEMError) as ei:
        ops.energy_raw(geom, u.to(DEV), f=b, load_vector=True)
    assert ei.value.code == L.DN_ENOSTREAM
    lref, gref = oracle_energy(fem, u, **kw)
    l1, g1 = _run(fem, u, True, **kw)
    assert_parity(l1, g1, lref, gref, what="fallback")


def test_load_vector_is_reassembled_when_f_gp_changes_in_place():
    fem, u, kw = _case(2, (64, 32), 2, 1, 2, False, 1, seed=9)
    geom = fem.geometry
    ud = u.to(DEV)
    f_gp = kw["f_gp"].to(DEV)
    dm = [(m.to(DEV), v) for m, v in kw["dirichlet"]]
    l_a, _ = ops.fem_energy_and_grad(geom, ud, f_gp=f_gp, dirichlet=dm, c_k=0.5)
    l_a2, _ = ops.fem_energy_and_grad(geom, ud, f_gp=f_gp, dirichlet=dm, c_k=0.5)      # cache hit
    assert float(l_a) == float(l_a2)
    prep = ops.PreparedEnergy(geom, ud, f_gp=f_gp, dirichlet=dm, c_k=0.5)
    assert float(prep()[0]) == float(l_a)
    f_gp.mul_(-2.0)                                                                      # bumps the version counter
    kw["f_gp"] = f_gp.cpu()
    lref, gref = oracle_energy(fem, u, **kw)
    l_b, g_b = ops.fem_energy_and_grad(geom, ud, f_gp=f_gp, dirichlet=dm, c_k=0.5)
    assert_parity(l_b, g_b, lref, gref, what="after in-place update")
    l_p, g_p = prep()
    assert_parity(l_p, g_p, lref, gref, what="prepared call after in-place update")


def test_residual_mode_is_not_offered_and_nodal_f_is_unchanged():
    """The flag lives in dn_consts: the residual entry points have none; a nodal f with flags = 0 is the old path."""
    fem, u, kw = _case(2, (64, 32), 2, 1, 2, True, 1, seed=3)
    f = torch.randn(2, 1, 32, 64)
    kw2 = dict(nu=kw["nu"], f=f, dirichlet=kw["dirichlet"])
    lref, gref = oracle_energy(fem, u, **kw2)
    l1, g1 = _run(fem, u, True, **kw2)
    assert_parity(l1, g1, lref, gref, what="nodal f")
