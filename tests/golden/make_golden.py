"""Generate tests/golden/*.npz from the REAL reference (adityabalu/DiffNet).

Run in the build container, where ``/root/reference`` is mounted:

    python tests/golden/make_golden.py

The reference's ``DiffNet2DFEM`` / ``DiffNet3DFEM`` objects are constructed unmodified
(Lightning stubbed, see oracle/refload.py) and supply every ``gauss_pt_evaluation*`` call
and every table (gpw, h, Nvalues, dN_*_values); the ``loss()`` bodies -- which live in
example scripts that cannot be imported here (matplotlib / lightning / libconf at module
top) -- are the restated bodies of ``oracle/losses.py`` called WITH THE REFERENCE OBJECT as
``fem``.  Gradients come from torch autograd on the CPU, fp32; an fp64 run of the same
(``module.double()``) is stored beside it as the noise-floor witness.

The vectors are small on purpose: they pin conventions (axis order, Gauss-point/basis
numbering, mask precedence, scaling), not performance.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import losses as L            # noqa: E402
from oracle.refload import load_reference  # noqa: E402

ref = load_reference()
torch.manual_seed(20261018)
torch.set_num_threads(1)


def grad_of(fn, u):
    u = u.clone().requires_grad_(True)
    loss = fn(u)
    (g,) = torch.autograd.grad(loss, u)
    return loss.detach(), g


def both_precisions(make_fem, body, u, *args):
    fem32 = make_fem()
    l32, g32 = grad_of(lambda v: body(fem32, v, *args), u)
    fem64 = make_fem().double()
    args64 = [a.double() if torch.is_tensor(a) else a for a in args]
    l64, g64 = grad_of(lambda v: body(fem64, v, *args64), u.double())
    return dict(loss=l32.numpy(), grad=g32.numpy(), loss64=l64.numpy(), grad64=g64.numpy())


def save(name, **arrays):
    flat = {}
    for k, v in arrays.items():
        if isinstance(v, dict):
            for kk, vv in v.items():
                flat[f"{k}.{kk}"] = np.asarray(vv)
        else:
            flat[k] = v.numpy() if torch.is_tensor(v) else np.asarray(v)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **flat)
    print(f"{name}: {os.path.getsize(path)} bytes, keys={len(flat)}")


def masks2d(B, H, W):
    bc1 = torch.zeros(B, 1, H, W); bc1[..., 0] = 1      # klsum.py:19-24
    bc2 = torch.zeros(B, 1, H, W); bc2[..., -1] = 1
    return bc1, bc2


# ---------------------------------------------------------------- 2-D, 12 x 9, B = 2
def case_2d_rect():
    X, Y, B = 12, 9, 2
    mk = lambda: ref.DiffNet2DFEM(None, domain_sizes=(X, Y, 1), domain_lengths=(1.5, 1.0, 1.0),
                                  domain_size=X, domain_length=1.5)
    fem = mk()
    u = torch.randn(B, 1, Y, X)
    nu = torch.exp(0.5 * torch.randn(B, 1, Y, X))
    f = torch.randn(B, 1, Y, X)
    bc1, bc2 = masks2d(B, Y, X)
    bc1[0, 0, 3:5, 4:7] = 1          # interior Dirichlet patch
    bc2[1, 0, 0, :] = 1              # overlaps bc1 at a corner: bc2 wins
    inputs = torch.cat([nu, bc1, bc2], 1)
    out = dict(u=u, inputs=inputs, forcing=f, sizes=np.array([X, Y]), lengths=np.array([1.5, 1.0]),
               h=fem.h, hx=fem.hx, hy=fem.hy)
    um = torch.where(bc1 > 0.5, 1.0 + u * 0, u)
    out["gp.N"] = fem.gauss_pt_evaluation(u)
    out["gp.dx"] = fem.gauss_pt_evaluation_der_x(u)
    out["gp.dy"] = fem.gauss_pt_evaluation_der_y(u)
    out["E1"] = both_precisions(mk, L.body_klsum_energy, u, inputs, f)
    out["E2"] = both_precisions(mk, L.body_0_base, u, inputs, f)
    out["resmin"] = both_precisions(mk, L.body_klsum_resmin, u, inputs, f)
    # E3 with a nodal Dirichlet FIELD (e8_2d_poisson_mms.py:165,175) and f at Gauss points
    u_bc = torch.randn(1, 1, Y, X)
    f_gp = torch.randn(1, 4, Y - 1, X - 1)
    e3 = lambda fm, v, nu_, bc_, ubc_, fgp_: L.energy_loss(
        fm, v, nu=nu_, f_gp=fgp_, dirichlet=[(bc_, ubc_)], c_k=0.5, c_f=1.0)
    out["u_bc"], out["f_gp"] = u_bc, f_gp
    out["E3fgp"] = both_precisions(mk, e3, u, nu, bc2, u_bc, f_gp)
    save("ref_2d_rect", **out)


# ---------------------------------------------------------------- 2-D Neumann IBN, 17^2
def case_2d_neumann():
    N, B = 17, 3
    mk = lambda: ref.DiffNet2DFEM(None, domain_size=N)
    yy, xx = torch.meshgrid(torch.linspace(0, 1, N), torch.linspace(0, 1, N), indexing="ij")
    obj = (((xx - 0.5) ** 2 + (yy - 0.45) ** 2) < 0.04).float()[None, None].repeat(B, 1, 1, 1)
    nu = torch.ones(B, 1, N, N) + 0.3 * torch.rand(B, 1, N, N)
    bc2, bc3 = masks2d(B, N, N)
    inputs = torch.cat([nu, obj, bc2, bc3], 1)
    u = torch.rand(B, 1, N, N)
    f = torch.zeros(B, 1, N, N)
    save("ref_2d_neumann", u=u, inputs=inputs, forcing=f,
         E5=both_precisions(mk, L.body_ibn2d_neumann, u, inputs, f))


# ---------------------------------------------------------------- 2-D, ngp_1d = 3 and 4
def case_2d_ngp():
    N, B = 10, 2
    u = torch.randn(B, 1, N, N)
    nu = torch.exp(0.3 * torch.randn(B, 1, N, N))
    f = torch.randn(B, 1, N, N)
    bc1, bc2 = masks2d(B, N, N)
    inputs = torch.cat([nu, bc1, bc2], 1)
    out = dict(u=u, inputs=inputs, forcing=f)
    for ngp in (3, 4):
        mk = lambda: ref.DiffNet2DFEM(None, domain_size=N, ngp_1d=ngp)
        out[f"E1_ngp{ngp}"] = both_precisions(mk, L.body_klsum_energy, u, inputs, f)
        fem = mk()
        out[f"gpw_ngp{ngp}"] = fem.gpw
        out[f"gp.N_ngp{ngp}"] = fem.gauss_pt_evaluation(u)
        out[f"gp.dx_ngp{ngp}"] = fem.gauss_pt_evaluation_der_x(u)
    save("ref_2d_ngp", **out)


# ---------------------------------------------------------------- 3-D, 7 x 6 x 5, B = 2
def case_3d_box():
    X, Y, Z, B = 7, 6, 5, 2
    mk = lambda: ref.DiffNet3DFEM(None, nsd=3, domain_sizes=(X, Y, Z),
                                  domain_lengths=(1.0, 0.8, 0.5), domain_size=X)
    fem = mk()
    u = torch.randn(B, 1, Z, Y, X)
    nu = torch.exp(0.5 * torch.randn(B, 1, Z, Y, X))
    f = torch.randn(B, 1, Z, Y, X)
    src = torch.zeros(B, 1, Z, Y, X); src[0, 0, 1:3, 2:4, 2:5] = 1; src[1, 0, 0, :, 0:3] = 1
    sink = torch.zeros(B, 1, Z, Y, X)
    sink[:, :, 0] = 1; sink[:, :, -1] = 1; sink[:, :, :, 0] = 1
    sink[:, :, :, -1] = 1; sink[..., 0] = 1; sink[..., -1] = 1        # IBN_3D.py:88-94
    inputs = torch.cat([nu, src, sink], 1)
    out = dict(u=u, inputs=inputs, forcing=f, source=src, sink=sink,
               sizes=np.array([X, Y, Z]), lengths=np.array([1.0, 0.8, 0.5]))
    out["gp.N"] = fem.gauss_pt_evaluation(u)
    out["gp.dx"] = fem.gauss_pt_evaluation_der_x(u)
    out["gp.dy"] = fem.gauss_pt_evaluation_der_y(u)
    out["gp.dz"] = fem.gauss_pt_evaluation_der_z(u)
    out["ibn3d"] = both_precisions(mk, L.body_ibn3d, u, src, sink, f)
    out["inobj"] = both_precisions(mk, L.body_solve_in_object, u, inputs, f)
    # bare (D,H,W) parameter, B = 1 inputs (solve_in_object_3d.py:198-199)
    out["inobj_bare"] = both_precisions(mk, L.body_solve_in_object, u[0, 0], inputs[:1], f[:1])
    save("ref_3d_box", **out)


# ---------------------------------------------------------------- tests/test.py, tests/test3D.py
def case_reference_tests():
    n2 = 14
    x = torch.linspace(0.0, 1.0, n2)
    xx, yy = torch.meshgrid(x, x, indexing="ij")            # tests/test.py:88 (legacy 'ij')
    u2 = (torch.sin(np.pi * xx) * torch.sin(np.pi * yy))[None, None].repeat(3, 1, 1, 1)
    k2 = torch.ones(3, 1, n2, n2) + 0.2 * torch.rand(3, 1, n2, n2)
    mk2 = lambda: ref.DiffNet2DFEM(None, domain_size=n2 + 2)
    r2 = both_precisions(mk2, lambda fm, v, k: L.body_test2d_residual(fm, v, k), u2, k2)
    n3 = 6
    z = torch.linspace(0.0, 1.0, n3)
    a, b, c = torch.meshgrid(z, z, z, indexing="ij")
    u3 = ((1 - a.permute(2, 1, 0)) ** 3)[None, None]        # tests/test3D.py:94-104
    k3 = torch.ones(1, 1, n3, n3, n3) + 0.2 * torch.rand(1, 1, n3, n3, n3)
    mk3 = lambda: ref.DiffNet3DFEM(None, domain_size=n3 + 2, nsd=3)
    r3 = both_precisions(mk3, lambda fm, v, k: L.body_test3d_residual(fm, v, k), u3, k3)
    save("ref_tests", u2=u2, k2=k2, res2d=r2, u3=u3, k3=k3, res3d=r3)


if __name__ == "__main__":
    case_2d_rect()
    case_2d_neumann()
    case_2d_ngp()
    case_3d_box()
    case_reference_tests()
