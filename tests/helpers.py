"""Shared helpers for the parity tests: oracle evaluation (CPU, fp64 truth) and tolerances.

Tolerances are the ones BASELINE.json's north_star states for fp32:
loss relative error <= 1e-5, gradient relative (L2) error <= 1e-4, masks/indexing bit-exact
(gradient exactly 0 on Dirichlet nodes).
"""
import torch

from conftest import rel_l2, rel_scalar
from oracle import losses as OL
from oracle.fem import Q1Oracle

LOSS_RTOL = 1e-5
GRAD_RTOL = 1e-4
GRAD_MAX_RTOL = 1e-4     # max |g - g_ref| / max |g_ref|: a handful of wrong seam nodes cannot hide in an L2 norm


def oracle_for(fem, dtype=torch.float64):
    """Q1Oracle with the same geometry as a diffnet_b200 module."""
    kw = dict(nsd=fem.nsd, domain_size=fem.domain_size, domain_length=fem.domain_length,
              domain_sizes=fem.domain_sizes_nd, domain_lengths=fem.domain_lengths_nd,
              ngp_1d=fem.ngp_1d, dtype=dtype)
    return Q1Oracle(**kw)


def to64(x):
    if torch.is_tensor(x):
        return x.detach().cpu().double()
    if isinstance(x, (list, tuple)):
        return type(x)(to64(v) for v in x)
    return x


def oracle_energy(fem, u, **kw):
    """(loss, grad) of oracle.losses.energy_loss in fp64 on the CPU."""
    o = oracle_for(fem)
    u64 = to64(u).requires_grad_(True)
    kw64 = {k: to64(v) for k, v in kw.items()}
    loss = OL.energy_loss(o, u64, **kw64)
    (g,) = torch.autograd.grad(loss, u64)
    return loss.detach(), g


def oracle_residual(fem, u, **kw):
    o = oracle_for(fem)
    u64 = to64(u).requires_grad_(True)
    kw64 = {k: to64(v) for k, v in kw.items()}
    loss = OL.residual_loss(o, u64, **kw64)
    (g,) = torch.autograd.grad(loss, u64)
    return loss.detach(), g


def assert_parity(loss, grad, loss_ref, grad_ref, masks=(), what=""):
    """loss/grad within the north_star tolerances; gradient EXACTLY zero on Dirichlet nodes."""
    loss, grad = torch.as_tensor(loss).detach().cpu(), torch.as_tensor(grad).detach().cpu()
    el, eg = rel_scalar(loss, loss_ref), rel_l2(grad, grad_ref.reshape(grad.shape))
    assert el <= LOSS_RTOL, f"{what}: loss rel err {el:.3e} (got {float(loss)}, want {float(loss_ref)})"
    assert eg <= GRAD_RTOL, f"{what}: grad rel-L2 err {eg:.3e}"
    gr = grad_ref.reshape(grad.shape).double()
    em = float((grad.double() - gr).abs().max() / gr.abs().max().clamp_min(1e-300))
    assert em <= GRAD_MAX_RTOL, f"{what}: grad max-norm err {em:.3e}"
    for m in masks:
        m = torch.as_tensor(m).detach().cpu()
        if grad.dim() == m.dim():
            hit = (m > 0.5).expand_as(grad)
            assert torch.count_nonzero(grad[hit]) == 0, f"{what}: gradient not exactly 0 on Dirichlet nodes"
    return el, eg
