"""Oracle (TEST INFRASTRUCTURE): the Poisson loss() bodies of the reference.

The reference keeps the integrand in user code: every example subclasses
DiffNet2DFEM/DiffNet3DFEM and writes a ``loss()``.  This file restates

* the whole family as one parametrised function, :func:`energy_loss`
  (variants E1..E6 and the f-at-Gauss-points form), and
  :func:`residual_loss` (the assembled-residual "resmin" form);
* a few bodies literally, one function per reference site, so that tests can
  show "the 3-line loss() on the fused op == the reference body":
    - :func:`body_0_base`            examples/poisson/single_instance/0_base.py:31-56
    - :func:`body_klsum_energy`      examples/poisson/single_instance/12_klsum.py:53-78
    - :func:`body_klsum_resmin`      examples/poisson/single_instance/12_klsum.py:80-132
    - :func:`body_ibn2d_neumann`     IBN/poisson-2d/parametric/e2_complex_immersed_background_neumann.py:33-60
    - :func:`body_ibn3d`             IBN/poisson-3d/parametric/IBN_3D.py:114-136
    - :func:`body_solve_in_object`   IBN/poisson-3d/non-parametric/solve_in_object_3d.py:75-102
    - :func:`body_test2d_residual`   tests/test.py:43-79   (ctor bug fixed)
    - :func:`body_test3d_residual`   tests/test3D.py:47-85

All of it is torch-on-CPU through ``oracle.fem.gp_eval`` (= F.conv2d/3d), with
autograd supplying the gradients, exactly as in the reference.
"""
from __future__ import annotations

import torch

from .fem import Q1Oracle


def _gpw_b(fem, like):
    """gpw as (1, ngp, 1, 1[, 1]) in the dtype of ``like`` -- the
    ``self.gpw.unsqueeze(-1)...unsqueeze(0).type_as(x)`` idiom of every reference loss().
    Works on a :class:`Q1Oracle` and on a real reference ``DiffNet2DFEM/3DFEM``."""
    return fem.gpw.reshape((1, -1) + (1,) * fem.nsd).to(like.dtype)


def _dirichlet(u, dirichlet):
    """``u = torch.where(mask > 0.5, value + u*0.0, u)`` applied in order (later
    entries win where masks overlap).  0_base.py:41-42, e8_2d_poisson_mms.py:165."""
    for mask, value in dirichlet:
        u = torch.where(mask > 0.5, value + u * 0.0, u)
    return u


def energy_density(fem: Q1Oracle, u, nu=None, f=None, f_gp=None, dirichlet=(),
                   nu_zero_mask=None, c_k=1.0, c_f=1.0, scale=1.0):
    """Per-element energy (B, *elems): scale * sum_g gpw_g * (c_k * nu_g * |grad u|_g^2 -
    c_f * u_g * f_g) -- ``res_elmwise`` after ``torch.sum(res_elmwise, 1)`` in every reference
    loss() (e.g. 0_base.py:51-54)."""
    u = _dirichlet(u, dirichlet)
    if nu is not None and nu_zero_mask is not None:      # e2_..._neumann.py:44
        nu = torch.where(nu_zero_mask > 0.5, nu * 0.0, nu)
    u_gp = fem.gauss_pt_evaluation(u)
    gsq = fem.gauss_pt_evaluation_der_x(u) ** 2 + fem.gauss_pt_evaluation_der_y(u) ** 2
    if fem.nsd == 3:
        gsq = gsq + fem.gauss_pt_evaluation_der_z(u) ** 2
    stiff = gsq if nu is None else fem.gauss_pt_evaluation(nu) * gsq
    if f_gp is None:
        f_gp = fem.gauss_pt_evaluation(f) if f is not None else None
    integrand = c_k * stiff
    if f_gp is not None and c_f != 0.0:
        integrand = integrand - c_f * (u_gp * f_gp)
    res = scale * _gpw_b(fem, u_gp) * integrand
    return torch.sum(res, 1)


def energy_loss(fem: Q1Oracle, u, nu=None, f=None, f_gp=None, dirichlet=(),
                nu_zero_mask=None, c_k=1.0, c_f=1.0, scale=1.0, reduction="mean"):
    """scale * gpw_g * (c_k * nu_g * |grad u|_g^2 - c_f * u_g * f_g), summed over
    Gauss points, then mean (or sum) over batch x elements.  SURVEY.md App. A.4."""
    res = energy_density(fem, u, nu, f, f_gp, dirichlet, nu_zero_mask, c_k, c_f, scale)
    return torch.mean(res) if reduction == "mean" else torch.sum(res)


def _q1_assemble(R, R_split, nsd):
    """Scatter-add of the per-element, per-basis residual into the nodes,
    12_klsum.py:46-51 (2-D) / tests/test3D.py:36-45 (3-D).  Local index
    a = 2*jbf + ibf (2-D), 4*kbf + 2*jbf + ibf (3-D)."""
    lo, hi = slice(0, -1), slice(1, None)
    if nsd == 2:
        for a, (sj, si) in enumerate([(lo, lo), (lo, hi), (hi, lo), (hi, hi)]):
            R[:, 0, sj, si] += R_split[:, a]
    else:
        a = 0
        for sk in (lo, hi):
            for sj in (lo, hi):
                for si in (lo, hi):
                    R[:, 0, sk, sj, si] += R_split[:, a]
                    a += 1
    return R


def residual_vector(fem: Q1Oracle, u, nu=None, f=None, dirichlet=(), jac=1.0):
    """Assembled Galerkin residual R(B,1,nodes) = sum_e sum_g JxW_g (nu_g grad N_a . grad u_g
    - N_a f_g), zeroed at Dirichlet nodes.  12_klsum.py:80-129."""
    nsd = fem.nsd
    u = _dirichlet(u, dirichlet)
    JxW = (fem.gpw.to(u.dtype) * jac).reshape((1, 1, -1) + (1,) * nsd)
    ux = fem.gauss_pt_evaluation_der_x(u).unsqueeze(1)
    uy = fem.gauss_pt_evaluation_der_y(u).unsqueeze(1)
    flux = fem.dN_x_values.to(u.dtype) * ux + fem.dN_y_values.to(u.dtype) * uy
    if nsd == 3:
        flux = flux + fem.dN_z_values.to(u.dtype) * fem.gauss_pt_evaluation_der_z(u).unsqueeze(1)
    if nu is not None:
        flux = fem.gauss_pt_evaluation(nu).unsqueeze(1) * flux
    R_split = torch.sum(flux * JxW, 2)
    if f is not None:
        rhs = fem.Nvalues.to(u.dtype) * fem.gauss_pt_evaluation(f).unsqueeze(1) * JxW
        R_split = R_split - torch.sum(rhs, 2)
    shape = (R_split.shape[0], 1) + tuple(s + 1 for s in R_split.shape[2:])
    R = _q1_assemble(torch.zeros(shape, dtype=u.dtype), R_split, nsd)
    for mask, _ in dirichlet:
        R = torch.where(mask > 0.5, R * 0.0, R)
    return R


def residual_loss(fem, u, nu=None, f=None, dirichlet=(), jac=1.0):
    """``loss = sum(R**2)``, 12_klsum.py:131."""
    return torch.sum(residual_vector(fem, u, nu, f, dirichlet, jac) ** 2)


# --------------------------------------------------------------------------
# literal bodies (one per reference site)
# --------------------------------------------------------------------------

def body_0_base(fem, u, inputs, forcing):
    """0_base.py:31-56 -- E2: 0.5*(h/2)^2 * gpw * (nu |grad u|^2 - u f), mean."""
    nu, bc1, bc2 = inputs[:, 0:1], inputs[:, 1:2], inputs[:, 2:3]
    u = torch.where(bc1 > 0.5, 1.0 + u * 0.0, u)
    u = torch.where(bc2 > 0.5, u * 0.0, u)
    nu_gp = fem.gauss_pt_evaluation(nu)
    f_gp = fem.gauss_pt_evaluation(forcing)
    u_gp = fem.gauss_pt_evaluation(u)
    u_x = fem.gauss_pt_evaluation_der_x(u)
    u_y = fem.gauss_pt_evaluation_der_y(u)
    jac = (0.5 * fem.h) ** 2 * _gpw_b(fem, nu_gp)
    res = 0.5 * jac * (nu_gp * (u_x ** 2 + u_y ** 2) - (u_gp * f_gp))
    return torch.mean(torch.sum(res, 1))


def body_klsum_energy(fem, u, inputs, forcing):
    """12_klsum.py:53-78 (same body: e1_complex_immersed_background.py:33-58) -- E1."""
    nu, bc1, bc2 = inputs[:, 0:1], inputs[:, 1:2], inputs[:, 2:3]
    u = torch.where(bc1 > 0.5, 1.0 + u * 0.0, u)
    u = torch.where(bc2 > 0.5, u * 0.0, u)
    nu_gp = fem.gauss_pt_evaluation(nu)
    f_gp = fem.gauss_pt_evaluation(forcing)
    u_gp = fem.gauss_pt_evaluation(u)
    u_x = fem.gauss_pt_evaluation_der_x(u)
    u_y = fem.gauss_pt_evaluation_der_y(u)
    res = _gpw_b(fem, nu_gp) * (nu_gp * (u_x ** 2 + u_y ** 2) - (u_gp * f_gp))
    return torch.mean(torch.sum(res, 1))


def body_klsum_resmin(fem, u, inputs, forcing):
    """12_klsum.py:80-132 -- assembled residual, trnsfrm_jac = 1.0, loss = sum(R^2)."""
    nu, bc1, bc2 = inputs[:, 0:1], inputs[:, 1:2], inputs[:, 2:3]
    N = fem.Nvalues.to(u.dtype)
    dNx, dNy = fem.dN_x_values.to(u.dtype), fem.dN_y_values.to(u.dtype)
    JxW = (fem.gpw.to(u.dtype) * 1.0)[None, None, :, None, None]
    u = torch.where(bc1 > 0.5, 1.0 + u * 0.0, u)
    u = torch.where(bc2 > 0.5, u * 0.0, u)
    nu_gp = fem.gauss_pt_evaluation(nu).unsqueeze(1)
    f_gp = fem.gauss_pt_evaluation(forcing).unsqueeze(1)
    u_x = fem.gauss_pt_evaluation_der_x(u).unsqueeze(1)
    u_y = fem.gauss_pt_evaluation_der_y(u).unsqueeze(1)
    lhs = nu_gp * (dNx * u_x + dNy * u_y) * JxW
    rhs = N * f_gp * JxW
    R_split = torch.sum(lhs, 2) - torch.sum(rhs, 2)
    R = _q1_assemble(torch.zeros_like(u), R_split, 2)
    R = torch.where(bc1 > 0.5, R * 0.0, R)
    R = torch.where(bc2 > 0.5, R * 0.0, R)
    return torch.sum(R ** 2)


def body_ibn2d_neumann(fem, u, inputs, forcing):
    """e2_complex_immersed_background_neumann.py:33-60 -- E5 (nu zeroed in the object)."""
    nu, bc1, bc2, bc3 = (inputs[:, i:i + 1] for i in range(4))
    nu = torch.where(bc1 > 0.5, nu * 0.0, nu)
    u = torch.where(bc2 > 0.5, 1.0 + u * 0.0, u)
    u = torch.where(bc3 > 0.5, u * 0.0, u)
    nu_gp = fem.gauss_pt_evaluation(nu)
    f_gp = fem.gauss_pt_evaluation(forcing)
    u_gp = fem.gauss_pt_evaluation(u)
    u_x = fem.gauss_pt_evaluation_der_x(u)
    u_y = fem.gauss_pt_evaluation_der_y(u)
    res = _gpw_b(fem, nu_gp) * (nu_gp * (u_x ** 2 + u_y ** 2) - (u_gp * f_gp))
    return torch.mean(torch.sum(res, 1))


def body_ibn3d(fem, u, source, sink, forcing):
    """IBN_3D.py:114-136 -- E4 (nu == 1), source/sink with the overlap fix :119-122."""
    u = torch.where(source > 0.5, 1.0 + (u * 0.0), u)
    source = torch.where(source > 0.5, 1.0 + (source * 0.0), source * 0.0)
    sink = torch.where(source == sink, sink * 0.0, sink)
    u = torch.where(sink > 0.5, u * 0.0, u)
    f_gp = fem.gauss_pt_evaluation(forcing)
    u_gp = fem.gauss_pt_evaluation(u)
    u_x = fem.gauss_pt_evaluation_der_x(u)
    u_y = fem.gauss_pt_evaluation_der_y(u)
    u_z = fem.gauss_pt_evaluation_der_z(u)
    res = _gpw_b(fem, u_gp) * (1.0 * (u_x ** 2 + u_y ** 2 + u_z ** 2) - (u_gp * f_gp))
    return torch.mean(torch.sum(res, 1))


def body_solve_in_object(fem, u, inputs, forcing):
    """solve_in_object_3d.py:75-102 -- E3 in 3-D; u may be a bare (D,H,W) parameter
    (``where`` broadcasting lifts it to 5-D, :85)."""
    nu, bc1 = inputs[:, 0:1], inputs[:, 1:2]
    u = torch.where(bc1 > 0.5, u * 0.0, u)
    nu_gp = fem.gauss_pt_evaluation(nu)
    f_gp = fem.gauss_pt_evaluation(forcing)
    u_gp = fem.gauss_pt_evaluation(u)
    u_x = fem.gauss_pt_evaluation_der_x(u)
    u_y = fem.gauss_pt_evaluation_der_y(u)
    u_z = fem.gauss_pt_evaluation_der_z(u)
    res = _gpw_b(fem, nu_gp) * (0.5 * nu_gp * (u_x ** 2 + u_y ** 2 + u_z ** 2) - u_gp * f_gp)
    return torch.mean(torch.sum(res, 1))


def _pad_for_test(fem, out, inp):
    """tests/test.py:49-52 / tests/test3D.py:54-57: replicate-pad nu by one node on
    every side; replicate-pad u in y (and z), then x-pad with 1 on the left and 0 on
    the right."""
    P = torch.nn.functional.pad
    n = fem.nsd
    inp_pad = P(inp, (1, 1) * n, "replicate")
    out_pad = P(out, (0, 0) + (1, 1) * (n - 1), "replicate")
    out_pad = P(out_pad, (1, 0) + (0, 0) * (n - 1), "constant", value=1)
    out_pad = P(out_pad, (0, 1) + (0, 0) * (n - 1), "constant", value=0)
    return out_pad, inp_pad


def body_test2d_residual(fem, output, inp):
    """tests/test.py:43-79: mean over batch of sum(R^2), R = assembled nu grad N . grad u
    with JxW = gpw * (h/2)^2.  ``fem`` must be built for the padded size."""
    u, nu = _pad_for_test(fem, output, inp)
    JxW = (fem.gpw.to(u.dtype) * (0.5 * fem.h) ** 2)[None, :, None, None]
    u_x = fem.gauss_pt_evaluation_der_x(u).unsqueeze(1)
    u_y = fem.gauss_pt_evaluation_der_y(u).unsqueeze(1)
    nu_gp = fem.gauss_pt_evaluation(nu).unsqueeze(1)
    vxux = fem.dN_x_values.to(u.dtype) * u_x * JxW
    vyuy = fem.dN_y_values.to(u.dtype) * u_y * JxW
    R_split = torch.sum(nu_gp * (vxux + vyuy), 2)
    R = _q1_assemble(torch.zeros_like(nu), R_split, 2)
    return torch.mean(torch.sum(R ** 2, (-1, -2, -3)))


def body_test3d_residual(fem, output, inp):
    """tests/test3D.py:47-85, JxW = gpw * (h/2)^3."""
    u, nu = _pad_for_test(fem, output, inp)
    JxW = (fem.gpw.to(u.dtype) * (0.5 * fem.h) ** 3)[None, :, None, None, None]
    u_x = fem.gauss_pt_evaluation_der_x(u).unsqueeze(1)
    u_y = fem.gauss_pt_evaluation_der_y(u).unsqueeze(1)
    u_z = fem.gauss_pt_evaluation_der_z(u).unsqueeze(1)
    nu_gp = fem.gauss_pt_evaluation(nu).unsqueeze(1)
    flux = (fem.dN_x_values.to(u.dtype) * u_x * JxW + fem.dN_y_values.to(u.dtype) * u_y * JxW
            + fem.dN_z_values.to(u.dtype) * u_z * JxW)
    R_split = torch.sum(nu_gp * flux, 2)
    R = _q1_assemble(torch.zeros_like(nu), R_split, 3)
    return torch.mean(torch.sum(R ** 2, (-1, -2, -3, -4)))
