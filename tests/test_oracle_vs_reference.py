"""The oracle restatement vs. the REAL reference module (bit-for-bit).

Runs only where /root/reference is mounted (the build container); on the GPU box the
same pinning is carried by tests/golden/*.npz (tests/test_oracle_golden.py).
"""
import pytest
import torch

from oracle import losses as L
from oracle.fem import Q1Oracle
from oracle.refload import load_reference, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="/root/reference not mounted")


@pytest.fixture(scope="module")
def ref():
    return load_reference()


def _same_tables(o, r, names):
    for n in names:
        a, b = getattr(o, n), getattr(r, n)
        if isinstance(a, list):
            assert len(a) == len(b)
            for x, y in zip(a, b):
                assert x.shape == y.shape and torch.equal(x, y.detach()), n
        else:
            assert torch.equal(torch.as_tensor(a), torch.as_tensor(b)), n


@pytest.mark.parametrize("ngp", [2, 3, 4])
def test_tables_2d_bit_exact(ref, ngp):
    kw = dict(domain_sizes=(12, 9, 1), domain_lengths=(1.5, 1.0, 1.0), domain_size=12,
              domain_length=1.5, ngp_1d=ngp)
    r = ref.DiffNet2DFEM(None, **kw)
    o = Q1Oracle(nsd=2, **kw)
    _same_tables(o, r, ["gpw", "N_gp", "dN_x_gp", "dN_y_gp", "Nvalues", "dN_x_values",
                        "dN_y_values", "xx", "yy", "xgp", "ygp"])
    assert o.h == r.h and o.hs[0] == r.hx and o.hs[1] == r.hy
    assert o.ngp_total == r.ngp_total and o.nelems == (r.nelemX, r.nelemY)


@pytest.mark.parametrize("ngp", [2, 3])
def test_tables_3d_bit_exact(ref, ngp):
    kw = dict(domain_sizes=(7, 6, 5), domain_lengths=(1.0, 0.8, 0.5), domain_size=7, ngp_1d=ngp)
    r = ref.DiffNet3DFEM(None, nsd=3, **kw)
    o = Q1Oracle(nsd=3, **kw)
    _same_tables(o, r, ["gpw", "N_gp", "dN_x_gp", "dN_y_gp", "dN_z_gp", "Nvalues", "dN_x_values",
                        "dN_y_values", "dN_z_values", "xx", "yy", "zz", "xgp", "ygp", "zgp"])
    assert (o.hs[0], o.hs[1], o.hs[2]) == (r.hx, r.hy, r.hz)


def test_default_kwargs_match(ref):
    r = ref.DiffNet2DFEM(None)
    o = Q1Oracle(nsd=2)
    assert o.sizes == (r.domain_sizeX, r.domain_sizeY) == (64, 64)
    assert o.h == r.h and torch.equal(o.N_gp[0], r.N_gp[0].detach())


def test_gp_eval_and_losses_bit_exact(ref):
    torch.manual_seed(3)
    r = ref.DiffNet2DFEM(None, domain_size=20)
    o = Q1Oracle(nsd=2, domain_size=20)
    u = torch.randn(3, 1, 20, 20)
    for m in ("gauss_pt_evaluation", "gauss_pt_evaluation_der_x", "gauss_pt_evaluation_der_y"):
        assert torch.equal(getattr(o, m)(u), getattr(r, m)(u))
    nu = torch.rand(3, 1, 20, 20) + 0.5
    bc1 = torch.zeros_like(nu); bc1[..., 0] = 1
    bc2 = torch.zeros_like(nu); bc2[..., -1] = 1
    inputs = torch.cat([nu, bc1, bc2], 1)
    f = torch.randn_like(nu)
    for body in (L.body_0_base, L.body_klsum_energy, L.body_klsum_resmin):
        ur = u.clone().requires_grad_(True)
        uo = u.clone().requires_grad_(True)
        lr, lo = body(r, ur, inputs, f), body(o, uo, inputs, f)
        lr.backward(); lo.backward()
        assert torch.equal(lr, lo) and torch.equal(ur.grad, uo.grad), body.__name__


def test_3d_losses_bit_exact(ref):
    torch.manual_seed(4)
    r = ref.DiffNet3DFEM(None, domain_size=9, nsd=3)
    o = Q1Oracle(nsd=3, domain_size=9)
    u = torch.randn(2, 1, 9, 9, 9)
    assert torch.equal(o.gauss_pt_evaluation_der_z(u), r.gauss_pt_evaluation_der_z(u))
    src = (torch.rand_like(u) > 0.9).float()
    sink = torch.zeros_like(u); sink[..., 0] = 1; sink[:, :, 0] = 1
    f = torch.randn_like(u)
    ur = u.clone().requires_grad_(True); uo = u.clone().requires_grad_(True)
    lr, lo = L.body_ibn3d(r, ur, src, sink, f), L.body_ibn3d(o, uo, src, sink, f)
    lr.backward(); lo.backward()
    assert torch.equal(lr, lo) and torch.equal(ur.grad, uo.grad)


def test_reference_checkpoints_load_strictly():
    """A state_dict written by the reference's DiffNet2DFEM / DiffNet3DFEM / UNet loads into the
    diffnet_b200 classes with strict=True (same keys, same shapes, same table values), and back."""
    import importlib
    from diffnet_b200 import DiffNet2DFEM, DiffNet3DFEM
    from diffnet_b200.networks import UNet
    ref = load_reference()
    for ours, theirs in ((DiffNet2DFEM(None, domain_size=16), ref.DiffNet2DFEM(None, domain_size=16)),
                         (DiffNet3DFEM(None, domain_size=8), ref.DiffNet3DFEM(None, domain_size=8, nsd=3))):
        sd = theirs.state_dict()
        assert set(sd) == set(ours.state_dict())
        for k, v in ours.state_dict().items():
            assert torch.equal(v, sd[k]), k
        ours.load_state_dict(sd, strict=True)
        theirs.load_state_dict(ours.state_dict(), strict=True)
    unets = importlib.import_module("DiffNet.networks.unets")
    torch.manual_seed(0)
    rnet, net = unets.UNet(3, 1).eval(), UNet(3, 1).eval()
    net.load_state_dict(rnet.state_dict(), strict=True)
    x = torch.rand(2, 3, 64, 64)
    with torch.no_grad():
        assert torch.equal(net(x), rnet(x))
    rnet.load_state_dict(net.reference_state_dict(), strict=True)
