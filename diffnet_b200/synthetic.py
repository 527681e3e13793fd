"""Synthetic inputs with the shapes/statistics of the reference's datasets (bench + tests).

Not on the hot path.  Recipes restated from the reference:
  * KL log-normal diffusivity nu = exp(sum_i a_i sqrt(l_x,i l_y,i) phi_i(x) phi_i(y)), 6 modes,
    eta = 0.5, a ~ U[-3,3]^6   (DiffNet/gen_input_calc.py:74-91,132-181; Sobol range of
    examples/poisson/parametric/sobol_6d.npy).  The frequencies omega_i are the roots of
    (eta^2 w^2 - 1) sin w = 2 eta w cos w, found here by bisection instead of a constant table.
  * bc1 = first column, bc2 = last column (DiffNet/datasets/parametric/klsum.py:24-32).
  * immersed-geometry image inputs [domain, bc1 = object, bc2 = the four edges]
    (DiffNet/datasets/parametric/images.py:9-49, ImageIMBack) with star-shaped random silhouettes
    standing in for the 1464 PNGs of IBN/datasets/imagedataset.tar.gz.
  * 3-D source/sink masks: union of random boxes as "source", the six faces as "sink"
    (IBN/poisson-3d/parametric/IBN_3D.py:76-104).
"""
from __future__ import annotations

import functools
import math

import numpy as np
import torch


def kl_omegas(eta: float, n: int = 6) -> np.ndarray:
    """First n positive roots of (eta^2 w^2 - 1) sin(w) - 2 eta w cos(w) = 0 (cached per (eta, n))."""
    return np.array(_kl_omegas(float(eta), int(n)))


@functools.lru_cache(maxsize=32)
def _kl_omegas(eta: float, n: int):
    g = lambda w: (eta * eta * w * w - 1.0) * math.sin(w) - 2.0 * eta * w * math.cos(w)
    roots, w, step = [], 1e-6, 1e-3
    prev = g(w)
    while len(roots) < n:
        w2 = w + step
        cur = g(w2)
        if prev * cur < 0:
            a, b = w, w2
            for _ in range(80):
                m = 0.5 * (a + b)
                if g(a) * g(m) <= 0:
                    b = m
                else:
                    a = m
            roots.append(0.5 * (a + b))
        w, prev = w2, cur
    return tuple(roots)


def kl_diffusivity_2d(coeffs: torch.Tensor, size: int, eta: float = 0.5) -> torch.Tensor:
    """coeffs (B,6) -> nu (B,1,size,size) float32 on coeffs.device."""
    dev = coeffs.device
    om = torch.tensor(kl_omegas(eta, coeffs.shape[1]), dtype=torch.float64, device=dev)
    lam = 2.0 * eta / (1.0 + (eta * om) ** 2)
    x = torch.linspace(0, 1, size, dtype=torch.float64, device=dev)
    phi = eta * om[:, None] * torch.cos(om[:, None] * x[None]) + torch.sin(om[:, None] * x[None])  # (6,size)
    amp = coeffs.double() * lam                      # sqrt(lx)*sqrt(ly) = lam for eta_x = eta_y
    field = torch.einsum("bi,iy,ix->byx", amp, phi, phi)
    return torch.exp(field).float().unsqueeze(1)


def poisson2d_parametric_batch(B: int, size: int, device, seed: int = 1234):
    """(u, inputs (B,3,H,W) = [nu, bc1, bc2], forcing (B,1,H,W)) like KLSumStochastic."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    coeffs = (torch.rand(B, 6, generator=g) * 6.0 - 3.0).to(device)
    nu = kl_diffusivity_2d(coeffs, size)
    bc1 = torch.zeros(B, 1, size, size, device=device); bc1[..., 0] = 1
    bc2 = torch.zeros(B, 1, size, size, device=device); bc2[..., -1] = 1
    inputs = torch.cat([nu, bc1, bc2], 1).contiguous()
    forcing = torch.zeros(B, 1, size, size, device=device)
    u = (torch.randn(B, 1, size, size, generator=g) * 0.5 + 0.5).to(device)
    return u, inputs, forcing


def star_params(B: int, seed: int = 1234):
    """(B, 11) float32 = (cx, cy, r0, a[4], phase[4]) of the synthetic star-shaped silhouettes
    r(theta) = r0 (1 + sum_k a_k cos(k theta + p_k)), k = 2..5, and the generator (for further draws)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    P = torch.zeros(B, 11)
    for b in range(B):
        c = (torch.rand(2, generator=g) - 0.5) * 0.6
        P[b, 0], P[b, 1] = c[0], c[1]
        P[b, 2] = 0.25 + 0.2 * float(torch.rand(1, generator=g))
        for k in range(2, 6):
            P[b, 3 + k - 2] = 0.25 * float(torch.rand(1, generator=g)) / k
            P[b, 7 + k - 2] = 6.2831853 * float(torch.rand(1, generator=g))
    return P, g


def _star_raster_cpu(P: torch.Tensor, size: int) -> torch.Tensor:
    """The silhouettes of `star_params` rasterised with torch on the CPU (the oracle legs of the bench)."""
    y, x = torch.meshgrid(torch.linspace(-1, 1, size), torch.linspace(-1, 1, size), indexing="ij")
    obj = torch.zeros(P.shape[0], 1, size, size)
    for b in range(P.shape[0]):
        cx, cy, r0 = float(P[b, 0]), float(P[b, 1]), float(P[b, 2])
        th = torch.atan2(y - cy, x - cx)
        rad = torch.full_like(th, r0)
        for k in range(2, 6):
            rad = rad + r0 * float(P[b, 3 + k - 2]) * torch.cos(k * th + float(P[b, 7 + k - 2]))
        obj[b, 0] = (torch.sqrt((x - cx) ** 2 + (y - cy) ** 2) < rad).float()
    return obj


def ibn2d_batch(B: int, size: int, device, seed: int = 1234):
    """(u, inputs (B,3,H,W) = [domain, bc1, bc2], forcing) like ImageIMBack: a random star-shaped
    object per sample, domain = 1 outside it, bc1 = the object (u -> 1), bc2 = the four edges (u -> 0).
    On a CUDA device the three channels are rasterised by the producer kernel (dn_gen_star_inputs_f32):
    44 bytes of parameters per sample cross PCIe instead of 12 bytes per node."""
    P, g = star_params(B, seed)
    if torch.device(device).type == "cuda":
        from .datasets import star_inputs
        inputs, forcing = star_inputs(P, size, device)
    else:
        obj = _star_raster_cpu(P, size)
        bc2 = torch.zeros(B, 1, size, size)
        bc2[..., 0] = 1; bc2[..., -1] = 1; bc2[:, :, 0, :] = 1; bc2[:, :, -1, :] = 1
        inputs = torch.cat([1.0 - obj, obj, bc2], 1).contiguous()
        forcing = torch.zeros(B, 1, size, size)
    u = (torch.randn(B, 1, size, size, generator=g) * 0.5 + 0.5).to(device)
    return u, inputs, forcing


def box_params(B: int, size: int, seed: int = 1234):
    """(B, 19) int32 = (n, lo[3][3], hi[3][3]): 1-3 random axis-aligned boxes per sample, and the generator."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    P = torch.zeros(B, 19, dtype=torch.int32)
    for b in range(B):
        n = int(torch.randint(1, 4, (1,), generator=g))
        P[b, 0] = n
        for q in range(n):
            lo = torch.randint(1, size // 2, (3,), generator=g)
            ext = torch.randint(size // 8, size // 3, (3,), generator=g)
            hi = torch.minimum(lo + ext, torch.tensor(size - 1))
            P[b, 1 + 3 * q:4 + 3 * q] = lo.int()
            P[b, 10 + 3 * q:13 + 3 * q] = hi.int()
    return P, g


def poisson3d_parametric_batch(B: int, size: int, device, seed: int = 1234):
    """(u, source, sink, forcing), each (B,1,D,H,W): source = union of 1-3 random boxes, sink = the six faces
    (stand-in for the SIMP topologies of IBN/poisson-3d/parametric/IBN_3D.py:76-104).  On a CUDA device the
    masks are written by the producer kernel (dn_gen_box_masks_3d_f32)."""
    P, g = box_params(B, size, seed)
    if torch.device(device).type == "cuda":
        from .datasets import box_masks_3d
        src, sink, f = box_masks_3d(P, size, device)
    else:
        src = torch.zeros(B, 1, size, size, size)
        for b in range(B):
            for q in range(int(P[b, 0])):
                lo, hi = P[b, 1 + 3 * q:4 + 3 * q], P[b, 10 + 3 * q:13 + 3 * q]
                src[b, 0, lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] = 1
        sink = torch.zeros(B, 1, size, size, size)
        sink[:, :, 0] = 1; sink[:, :, -1] = 1; sink[:, :, :, 0] = 1
        sink[:, :, :, -1] = 1; sink[..., 0] = 1; sink[..., -1] = 1
        f = torch.zeros(B, 1, size, size, size)
    u = torch.rand(B, 1, size, size, size, generator=g).to(device)
    return u, src, sink, f
