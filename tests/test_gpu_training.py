"""The callers of the hot path on the GPU: reference-style modules (diffnet_b200/poisson.py)
driven by the minimal trainer -- the fused loss inside a real optimisation loop."""
import pytest
import torch

from diffnet_b200.networks import UNet
from diffnet_b200.poisson import PoissonInObject3D, PoissonParametric2D
from diffnet_b200.synthetic import poisson2d_parametric_batch
from diffnet_b200.trainer import Trainer

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def test_parametric_2d_training_reduces_the_energy():
    """UNet(3,1) + PoissonParametric2D on 64^2, B = 8 (2_klsum_fem.py style): Adam steps through
    the fused op lower the Galerkin energy; gradients reach every network parameter."""
    torch.manual_seed(0)
    net = UNet(3, 1, dropout=False)
    mod = PoissonParametric2D(net, domain_size=64, batch_size=8, learning_rate=1e-3)
    _, inputs, f = poisson2d_parametric_batch(8, 64, DEV, seed=5)
    tr = Trainer(max_steps=30, device=DEV, ddp=False)
    tr.fit(mod, [(inputs, f)] * 30)
    losses = torch.stack(tr.losses).cpu()
    assert torch.isfinite(losses).all()
    assert float(losses[-5:].mean()) < float(losses[:5].mean())
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in net.parameters())


def test_nonparametric_3d_u_as_parameter():
    """solve_in_object_3d.py: u is the parameter (a bare (D,H,W) tensor in a ParameterList), Adam on it;
    the energy decreases and Dirichlet nodes never move."""
    N = 24
    g = torch.Generator().manual_seed(1)
    zz, yy, xx = torch.meshgrid(*(torch.linspace(-0.5, 0.5, N),) * 3, indexing="ij")
    inside = ((xx ** 2 + yy ** 2 + zz ** 2) < 0.16).float()[None, None]
    inputs = torch.cat([inside, 1.0 - inside, torch.zeros_like(inside)], 1).to(DEV)
    forcing = torch.full((1, 1, N, N, N), 500.0, device=DEV)
    u0 = 0.01 * torch.randn(N, N, N, generator=g)
    net = torch.nn.ParameterList([torch.nn.Parameter(u0.clone())])
    mod = PoissonInObject3D(net, domain_size=N, learning_rate=1e-3)
    tr = Trainer(max_steps=40, device=DEV, ddp=False)
    tr.fit(mod, [(inputs, forcing)] * 40)
    losses = torch.stack(tr.losses).cpu()
    assert float(losses[-1]) < float(losses[0])
    outside = (inputs[0, 1] > 0.5).cpu()
    assert torch.equal(mod.network[0].detach().cpu()[outside], u0[outside])      # zero gradient there
