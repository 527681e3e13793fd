"""Oracle (TEST INFRASTRUCTURE): conv-based Gauss-point evaluation on Q1 meshes.

Restates reference ``DiffNet/DiffNetFEM.py``:
  * ``gauss_pt_eval``                      :7-18   -> :func:`gp_eval`
  * ``gauss_guadrature_scheme``            :128-141 -> :func:`gauss_rule`
  * linear 1-D basis (deg 1)               :54-59  -> :func:`q1_basis_1d`
  * element counts / h                     :42-51  -> :class:`Q1Oracle.__init__`
  * 2-D stencil tables, gpw, Nvalues, xgp  :178-235 -> :meth:`Q1Oracle._build`
  * 3-D stencil tables, gpw, xgp/ygp/zgp   :382-465 -> :meth:`Q1Oracle._build`
  * ``CuboidMesh.meshgrid_3d`` axis order  ``DiffNet/cuboid_mesh.py:8-25``

The stencil entries are computed in float64 in the same operation order as the
reference (``(b_x * b_y [* b_z]) * (2/h)``) and rounded to float32 on store, so
they are bit-identical to the reference's ``nn.Parameter`` tables (tested in
``tests/test_oracle_vs_reference.py``).

Only fem_basis_deg == 1 is restated: deg 2/3 crash upstream on numpy >= 1.24
(``np.float``), see SURVEY.md App. B.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def gauss_rule(ngp_1d: int):
    """1-D Gauss points/weights with the reference's (truncated) constants,
    DiffNetFEM.py:128-141."""
    if ngp_1d == 1:
        return np.array([0.0]), np.array([2.0])
    if ngp_1d == 2:
        return (np.array([-0.5773502691896258, 0.5773502691896258]),
                np.array([1.0, 1.0]))
    if ngp_1d == 3:
        return (np.array([-0.774596669, 0.0, +0.774596669]),
                np.array([5.0 / 9.0, 8.0 / 9.0, 5.0 / 9.0]))
    if ngp_1d == 4:
        return (np.array([-0.861136, -0.339981, +0.339981, +0.861136]),
                np.array([0.347855, 0.652145, 0.652145, 0.347855]))
    raise ValueError("ngp_1d must be 1..4")


def q1_basis_1d(x: float):
    """(values, derivatives) of the two linear Lagrange functions at x in [-1,1],
    DiffNetFEM.py:58-59."""
    return (np.array([0.5 * (1.0 - x), 0.5 * (1.0 + x)]),
            np.array([0.5 * (0.0 - 1.0), 0.5 * (0.0 + 1.0)]))


def gp_eval(tensor: torch.Tensor, stencils, nsd: int) -> torch.Tensor:
    """One cross-correlation per Gauss point, stride 1 (= nbf_1d - 1), no padding,
    concatenated on the channel axis.  DiffNetFEM.py:7-18."""
    conv = {1: F.conv1d, 2: F.conv2d, 3: F.conv3d}[nsd]
    return torch.cat([conv(tensor, w, stride=1) for w in stencils], dim=1)


class Q1Oracle:
    """Holds the tables ``DiffNet2DFEM`` / ``DiffNet3DFEM`` build at construction.

    domain_sizes / domain_lengths are (X, Y[, Z]) like the reference's
    ``domain_sizes`` kwarg (``DiffNet/base.py:23-32``); nodal tensors are
    (B,1,sizeY,sizeX) or (B,1,sizeZ,sizeY,sizeX), x contiguous.
    """

    def __init__(self, nsd=2, domain_size=64, domain_length=1.0,
                 domain_sizes=None, domain_lengths=None, ngp_1d=2,
                 dtype=torch.float32):
        assert nsd in (2, 3)
        self.nsd = nsd
        sizes = tuple(domain_sizes) if domain_sizes is not None else (domain_size,) * 3
        lengths = tuple(domain_lengths) if domain_lengths is not None else (domain_length,) * 3
        self.sizes = sizes[:nsd]            # (X, Y[, Z])
        self.lengths = lengths[:nsd]
        self.ngp_1d = max(int(ngp_1d), 2)   # DiffNetFEM.py:29-38 (deg 1 forces >= 2)
        self.nbf_1d = 2
        self.ngp_total = self.ngp_1d ** nsd
        self.nbf_total = self.nbf_1d ** nsd
        self.nelems = tuple(int(s - 1) for s in self.sizes)       # :42-46
        self.hs = tuple(l / n for l, n in zip(self.lengths, self.nelems))  # :47-51
        # backward-compat scalar: always from the domain_size/domain_length kwargs (:46,:51)
        self.h = domain_length / int(domain_size - 1)
        self.gpx_1d, self.gpw_1d = gauss_rule(self.ngp_1d)
        self.dtype = dtype
        self._build()

    # ------------------------------------------------------------------ tables
    def _build(self):
        nsd, ng, nb = self.nsd, self.ngp_1d, self.nbf_1d
        B = [q1_basis_1d(x)[0] for x in self.gpx_1d]   # B[gp][bf]
        D = [q1_basis_1d(x)[1] for x in self.gpx_1d]
        self.gpw = torch.zeros(self.ngp_total)
        self.N_gp, self.dN_x_gp, self.dN_y_gp, self.dN_z_gp = [], [], [], []
        tail = (1,) * nsd
        self.Nvalues = torch.ones((1, self.nbf_total, self.ngp_total) + tail)
        self.dN_x_values = torch.ones_like(self.Nvalues)
        self.dN_y_values = torch.ones_like(self.Nvalues)
        self.dN_z_values = torch.ones_like(self.Nvalues) if nsd == 3 else None
        sx = 2 / self.hs[0]
        sy = 2 / self.hs[1]
        sz = 2 / self.hs[2] if nsd == 3 else None

        def store(lst, arr64):
            t = torch.zeros(arr64.shape)          # float32, like the reference
            t.copy_(torch.from_numpy(arr64))      # f64 -> f32 rounding on store
            lst.append(t[None, None].clone())
            return t

        kgps = range(ng) if nsd == 3 else [None]
        for kgp in kgps:
            for jgp in range(ng):
                for igp in range(ng):
                    if nsd == 2:
                        G = ng * jgp + igp
                        self.gpw[G] = self.gpw_1d[igp] * self.gpw_1d[jgp]
                        # index [jbf, ibf]; products in reference order x * y
                        n = B[igp][None, :] * B[jgp][:, None]
                        dx = D[igp][None, :] * B[jgp][:, None] * sx
                        dy = B[igp][None, :] * D[jgp][:, None] * sy
                        tabs = [(self.N_gp, self.Nvalues, n),
                                (self.dN_x_gp, self.dN_x_values, dx),
                                (self.dN_y_gp, self.dN_y_values, dy)]
                    else:
                        G = kgp * ng * ng + jgp * ng + igp
                        self.gpw[G] = self.gpw_1d[igp] * self.gpw_1d[jgp] * self.gpw_1d[kgp]
                        bi, bj, bk = (B[igp][None, None, :], B[jgp][None, :, None],
                                      B[kgp][:, None, None])
                        di, dj, dk = (D[igp][None, None, :], D[jgp][None, :, None],
                                      D[kgp][:, None, None])
                        n = bi * bj * bk
                        dx = di * bj * bk * sx
                        dy = bi * dj * bk * sy
                        dz = bi * bj * dk * sz
                        tabs = [(self.N_gp, self.Nvalues, n),
                                (self.dN_x_gp, self.dN_x_values, dx),
                                (self.dN_y_gp, self.dN_y_values, dy),
                                (self.dN_z_gp, self.dN_z_values, dz)]
                    for lst, vals, arr in tabs:
                        t32 = store(lst, np.ascontiguousarray(arr))
                        vals[0, :, G] = t32.reshape((self.nbf_total,) + tail)

        for name in ("N_gp", "dN_x_gp", "dN_y_gp", "dN_z_gp"):
            setattr(self, name, [w.to(self.dtype) for w in getattr(self, name)])

        # nodal coordinates and their Gauss-point images (:229-235, :455-465)
        x = np.linspace(0, self.lengths[0], self.sizes[0])
        y = np.linspace(0, self.lengths[1], self.sizes[1])
        if nsd == 2:
            xx, yy = np.meshgrid(x, y)
            self.xx, self.yy = torch.FloatTensor(xx), torch.FloatTensor(yy)
        else:
            z = np.linspace(0, self.lengths[2], self.sizes[2])
            M, N, P = len(x), len(y), len(z)        # cuboid_mesh.py:8-25 -> (P,N,M)
            x2, y2 = np.meshgrid(x, y)
            self.xx = torch.FloatTensor(np.tile(x2, (P, 1, 1)))
            self.yy = torch.FloatTensor(np.tile(y2, (P, 1, 1)))
            self.zz = torch.FloatTensor(np.reshape(np.repeat(z, N * M), (P, N, M)))
        f32 = [w.float() for w in self.N_gp]
        self.xgp = gp_eval(self.xx[None, None], f32, nsd)
        self.ygp = gp_eval(self.yy[None, None], f32, nsd)
        if nsd == 3:
            self.zgp = gp_eval(self.zz[None, None], f32, nsd)

    # ------------------------------------------------- DiffNetFEM.py:143-156
    def gauss_pt_evaluation(self, t):
        return gp_eval(t, self.N_gp, self.nsd)

    def gauss_pt_evaluation_der_x(self, t):
        return gp_eval(t, self.dN_x_gp, self.nsd)

    def gauss_pt_evaluation_der_y(self, t):
        return gp_eval(t, self.dN_y_gp, self.nsd)

    def gauss_pt_evaluation_der_z(self, t):
        return gp_eval(t, self.dN_z_gp, self.nsd)


# ------------------------------------------------------------------ degree 2 / 3 and the surface stencils
def lagrange_basis_1d(deg: int):
    """(values(x), derivatives(x)) of the equispaced Lagrange basis of degree 1..3 on [-1,1],
    DiffNetFEM.py:58-59 (deg 1), :71-80 (deg 2), :105-117 (deg 3).  Upstream writes ``dtype=np.float`` in the
    deg 2/3 lambdas, which numpy >= 1.24 rejects; the polynomials themselves are restated unchanged."""
    if deg == 1:
        return (lambda x: np.array([0.5 * (1. - x), 0.5 * (1. + x)]),
                lambda x: np.array([0.5 * (0. - 1.), 0.5 * (0. + 1.)]))
    if deg == 2:
        return (lambda x: np.array([0.5 * x * (x - 1.), (1. - x ** 2), 0.5 * x * (x + 1.)], dtype=float),
                lambda x: np.array([0.5 * (2. * x - 1.), (- 2. * x), 0.5 * (2. * x + 1.)], dtype=float))
    if deg == 3:
        return (lambda x: np.array([(-9. / 16.) * (x ** 3 - x ** 2 - (1. / 9.) * x + (1. / 9.)),
                                    (27. / 16.) * (x ** 3 - (1. / 3.) * x ** 2 - x + (1. / 3.)),
                                    (-27. / 16.) * (x ** 3 + (1. / 3.) * x ** 2 - x - (1. / 3.)),
                                    (9. / 16.) * (x ** 3 + x ** 2 - (1. / 9.) * x - (1. / 9.))], dtype=float),
                lambda x: np.array([(-9. / 16.) * (3 * x ** 2 - 2 * x - (1. / 9.)),
                                    (27. / 16.) * (3 * x ** 2 - (2. / 3.) * x - 1),
                                    (-27. / 16.) * (3 * x ** 2 + (2. / 3.) * x - 1),
                                    (9. / 16.) * (3 * x ** 2 + 2 * x - (1. / 9.))], dtype=float))
    raise ValueError("fem_basis_deg must be 1, 2 or 3")


class LagrangeOracle:
    """gauss_pt_evaluation{,_der_x,_der_y,_der_z,_surf} for fem_basis_deg 1..3: one strided convolution per Gauss
    point (stride = nbf_1d - 1, DiffNetFEM.py:143-156 with :7-18), stencils built as :197-215 / :405-453 do, surface
    stencils as :244-269.  sizes / lengths are (X, Y[, Z])."""

    def __init__(self, nsd, sizes, lengths, fem_basis_deg=1, ngp_1d=2, dtype=torch.float64):
        self.nsd, self.deg = nsd, fem_basis_deg
        self.ngp_1d = max(int(ngp_1d), 2 if fem_basis_deg == 1 else 3)           # :27-38
        self.nbf_1d = nb = fem_basis_deg + 1
        self.gpx_1d, self.gpw_1d = gauss_rule(self.ngp_1d)
        nel = [int((s - 1) / fem_basis_deg) for s in sizes[:nsd]]                # :42-46
        hs = [l / n for l, n in zip(lengths[:nsd], nel)]
        bf, der = lagrange_basis_1d(fem_basis_deg)
        B = [bf(x) for x in self.gpx_1d]
        D = [der(x) for x in self.gpx_1d]
        ng = self.ngp_1d
        self.tables = {k: [] for k in ("N", "dx", "dy", "dz")[:nsd + 1]}
        for gp in np.ndindex(*(ng,) * nsd):                                      # ([kg,] jg, ig)
            g = gp[::-1]                                                         # (ig, jg[, kg])
            for w, key in enumerate(self.tables):
                fac = [D[g[d]] if w == d + 1 else B[g[d]] for d in range(nsd)]
                tab = fac[0]
                for d in range(1, nsd):
                    tab = tab * fac[d].reshape((nb,) + (1,) * d)                 # x * y [* z], like the reference
                if w > 0:
                    tab = tab * (2 / hs[w - 1])
                t32 = torch.zeros(tab.shape)
                t32.copy_(torch.from_numpy(np.ascontiguousarray(tab)))           # f64 -> f32 rounding on store
                self.tables[key].append(t32[None, None].to(dtype))
        self.surf = [torch.tensor(B[g], dtype=torch.float32)[None, None].to(dtype) for g in range(ng)]

    def _eval(self, t, key):
        conv = {2: F.conv2d, 3: F.conv3d}[self.nsd]
        return torch.cat([conv(t, w, stride=self.nbf_1d - 1) for w in self.tables[key]], dim=1)

    def gauss_pt_evaluation(self, t):
        return self._eval(t, "N")

    def gauss_pt_evaluation_der_x(self, t):
        return self._eval(t, "dx")

    def gauss_pt_evaluation_der_y(self, t):
        return self._eval(t, "dy")

    def gauss_pt_evaluation_der_z(self, t):
        return self._eval(t, "dz")

    def gauss_pt_evaluation_surf(self, t):
        return torch.cat([F.conv1d(t, w, stride=self.nbf_1d - 1) for w in self.surf], dim=1)
