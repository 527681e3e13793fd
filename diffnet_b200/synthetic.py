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

import math

import numpy as np
import torch


def kl_omegas(eta: float, n: int = 6) -> np.ndarray:
    """First n positive roots of (eta^2 w^2 - 1) sin(w) - 2 eta w cos(w) = 0."""
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
    return np.array(roots)


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


def ibn2d_batch(B: int, size: int, device, seed: int = 1234):
    """(u, inputs (B,3,H,W) = [domain, bc1, bc2], forcing) like ImageIMBack: a random star-shaped
    object per sample (radius r(theta) = r0 (1 + sum_k a_k cos(k theta + p_k))), domain = 1 outside
    it, bc1 = the object (u -> 1), bc2 = the four edges (u -> 0)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    y, x = torch.meshgrid(torch.linspace(-1, 1, size), torch.linspace(-1, 1, size), indexing="ij")
    obj = torch.zeros(B, 1, size, size)
    for b in range(B):
        c = (torch.rand(2, generator=g) - 0.5) * 0.6
        r0 = 0.25 + 0.2 * float(torch.rand(1, generator=g))
        th = torch.atan2(y - c[1], x - c[0])
        rad = torch.full_like(th, r0)
        for k in range(2, 6):
            a = 0.25 * float(torch.rand(1, generator=g)) / k
            ph = 6.2831853 * float(torch.rand(1, generator=g))
            rad = rad + r0 * a * torch.cos(k * th + ph)
        obj[b, 0] = (torch.sqrt((x - c[0]) ** 2 + (y - c[1]) ** 2) < rad).float()
    domain = 1.0 - obj
    bc2 = torch.zeros(B, 1, size, size)
    bc2[..., 0] = 1; bc2[..., -1] = 1; bc2[:, :, 0, :] = 1; bc2[:, :, -1, :] = 1
    inputs = torch.cat([domain, obj, bc2], 1).contiguous().to(device)
    forcing = torch.zeros(B, 1, size, size, device=device)
    u = (torch.randn(B, 1, size, size, generator=g) * 0.5 + 0.5).to(device)
    return u, inputs, forcing


def poisson3d_parametric_batch(B: int, size: int, device, seed: int = 1234):
    """(u, source, sink, forcing), each (B,1,D,H,W)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    src = torch.zeros(B, 1, size, size, size)
    for b in range(B):
        for _ in range(int(torch.randint(1, 4, (1,), generator=g))):
            lo = torch.randint(1, size // 2, (3,), generator=g)
            ext = torch.randint(size // 8, size // 3, (3,), generator=g)
            hi = torch.minimum(lo + ext, torch.tensor(size - 1))
            src[b, 0, lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] = 1
    sink = torch.zeros(B, 1, size, size, size)
    sink[:, :, 0] = 1; sink[:, :, -1] = 1; sink[:, :, :, 0] = 1
    sink[:, :, :, -1] = 1; sink[..., 0] = 1; sink[..., -1] = 1
    u = torch.rand(B, 1, size, size, size, generator=g)
    f = torch.zeros(B, 1, size, size, size)
    return u.to(device), src.to(device), sink.to(device), f.to(device)
