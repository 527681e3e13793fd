"""Host-side mirror of the reference module API (no GPU): tables, kwargs, attribute names."""
import numpy as np
import pytest
import torch

from diffnet_b200 import DiffNet2DFEM, DiffNet3DFEM, PDE
from oracle.fem import Q1Oracle


@pytest.mark.parametrize("ngp", [2, 3, 4])
def test_2d_tables_equal_oracle(ngp):
    kw = dict(domain_sizes=(12, 9, 1), domain_lengths=(1.5, 1.0, 1.0), domain_size=12,
              domain_length=1.5, ngp_1d=ngp)
    m, o = DiffNet2DFEM(None, **kw), Q1Oracle(nsd=2, **kw)
    for n in ("N_gp", "dN_x_gp", "dN_y_gp"):
        assert all(torch.equal(a.detach(), b) for a, b in zip(getattr(m, n), getattr(o, n))), n
    for n in ("gpw", "Nvalues", "dN_x_values", "dN_y_values", "xx", "yy"):
        assert torch.equal(getattr(m, n), getattr(o, n)), n
    assert torch.allclose(m.xgp, o.xgp, atol=2e-7) and torch.allclose(m.ygp, o.ygp, atol=2e-7)
    assert (m.h, m.hx, m.hy, m.nelemX, m.nelemY) == (o.h, o.hs[0], o.hs[1], 11, 8)
    assert m.ngp_total == ngp * ngp and m.nbf_total == 4


def test_3d_tables_equal_oracle():
    kw = dict(domain_sizes=(7, 6, 5), domain_lengths=(1.0, 0.8, 0.5), domain_size=7)
    m, o = DiffNet3DFEM(None, **kw), Q1Oracle(nsd=3, **kw)
    for n in ("N_gp", "dN_x_gp", "dN_y_gp", "dN_z_gp"):
        assert all(torch.equal(a.detach(), b) for a, b in zip(getattr(m, n), getattr(o, n))), n
    for n in ("gpw", "Nvalues", "dN_x_values", "dN_y_values", "dN_z_values", "xx", "yy", "zz"):
        assert torch.equal(getattr(m, n), getattr(o, n)), n
    for n in ("xgp", "ygp", "zgp"):
        assert torch.allclose(getattr(m, n), getattr(o, n), atol=2e-7), n


def test_reference_kwargs_and_state_dict_keys():
    m = DiffNet2DFEM(None)                       # defaults: base.py:16-24
    assert (m.nsd, m.domain_size, m.batch_size, m.learning_rate) == (2, 64, 64, 3e-4)
    assert isinstance(m, PDE)
    keys = set(m.state_dict().keys())            # stencils are Parameters in the reference too
    assert {"N_gp.0", "N_gp.3", "dN_x_gp.0", "dN_y_gp.3"} <= keys
    assert not any(p.requires_grad for p in m.parameters())
    # ngp_1d below 2 is lifted to 2 for Q1 (DiffNetFEM.py:29-38)
    assert DiffNet2DFEM(None, ngp_1d=1).ngp_1d == 2
    # degree 2 / 3 are built (general-basis ops); like the reference they need (domain_size - 1) % degree == 0
    with pytest.raises(AssertionError):
        DiffNet2DFEM(None, fem_basis_deg=2)                 # default domain_size 64: 63 % 2 != 0
    assert DiffNet2DFEM(None, fem_basis_deg=2, domain_size=65).nbf_1d == 3
    with pytest.raises(NotImplementedError):
        DiffNet2DFEM(None, fem_basis_deg=4)


def test_user_subclass_contract():
    """A reference-style subclass: ctor passes kwargs through, loss() is user code."""
    class Poisson(DiffNet2DFEM):
        def loss(self, u, inputs_tensor, forcing_tensor):
            return self.energy_loss(u, nu=inputs_tensor[:, 0:1], f=forcing_tensor,
                                    dirichlet=[(inputs_tensor[:, 1:2], 1.0), (inputs_tensor[:, 2:3], 0.0)])
    p = Poisson(torch.nn.Identity(), domain_size=16, batch_size=4, learning_rate=1e-2)
    opts, scheds = p.configure_optimizers() if list(p.network.parameters()) else ([None], [])
    assert p.geometry.nx == 16 and p.geometry.hx == pytest.approx(1 / 15)


def test_load_vector_cache_host_logic(monkeypatch):
    """ops._cached_load_vector (the f_gp -> assembled load vector memo, host side only; the assembly kernel is
    replaced by a counter): one assembly per tensor OBJECT and version, re-assembly after an in-place update, a new
    tensor with equal contents is a different key, dead tensors leave the cache, the cache is bounded."""
    from diffnet_b200 import ops
    fem = DiffNet2DFEM(None, domain_size=16)
    calls = []

    def fake(geom, f_gp, out=None):
        calls.append(f_gp._version)
        return torch.zeros((f_gp.shape[0],) + geom.spatial)
    monkeypatch.setattr(ops, "load_vector", fake)
    monkeypatch.setattr(ops, "_LV_CACHE", [])
    f = torch.randn(1, 4, 15, 15)
    b0 = ops._cached_load_vector(fem.geometry, f)
    assert ops._cached_load_vector(fem.geometry, f) is b0 and len(calls) == 1          # hit
    f.mul_(2.0)                                                                         # version counter moves
    b1 = ops._cached_load_vector(fem.geometry, f)
    assert b1 is not b0 and len(calls) == 2 and len(ops._LV_CACHE) == 1
    g = f.clone()                                                                       # equal contents, other object
    ops._cached_load_vector(fem.geometry, g)
    assert len(calls) == 3 and len(ops._LV_CACHE) == 2
    fem3 = DiffNet2DFEM(None, domain_size=16, ngp_1d=3)                                 # same tensor, other rule: other key
    ops._cached_load_vector(fem3.geometry, f)
    assert len(calls) == 4
    del g
    import gc
    gc.collect()
    keep = [torch.randn(1, 4, 15, 15) for _ in range(2 * ops._LV_CACHE_MAX)]
    for t in keep:
        ops._cached_load_vector(fem.geometry, t)
    assert len(ops._LV_CACHE) <= ops._LV_CACHE_MAX
    assert all(ref() is not None for ref, *_ in ops._LV_CACHE)
