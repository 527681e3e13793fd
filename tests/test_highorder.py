"""fem_basis_deg 2 / 3 and gauss_pt_evaluation_surf: the oracle's restatement (oracle/fem.py: LagrangeOracle) and the
general-basis CUDA ops against outputs of the REAL reference (tests/golden/highorder.npz, written by
tests/golden/make_golden_highorder.py with the np.float alias restored for the upstream lambdas)."""
import os

import numpy as np
import pytest
import torch

from conftest import ROOT, rel_l2
from oracle.fem import LagrangeOracle

G = np.load(os.path.join(ROOT, "tests", "golden", "highorder.npz"))
CASES = {
    "d2_2d": dict(nsd=2, sizes=(9, 7, 1), lengths=(1.2, 0.9, 1.0), deg=2, ngp=2, ctor=dict(domain_size=9, fem_basis_deg=2, domain_sizes=(9, 7, 1), domain_lengths=(1.2, 0.9, 1.0))),
    "d3_2d": dict(nsd=2, sizes=(10, 10, 1), lengths=(1.0, 1.0, 1.0), deg=3, ngp=4, ctor=dict(domain_size=10, fem_basis_deg=3, ngp_1d=4)),
    "d2_3d": dict(nsd=3, sizes=(7, 5, 5), lengths=(1.0, 0.8, 0.6), deg=2, ngp=2, ctor=dict(domain_size=7, fem_basis_deg=2, nsd=3, domain_sizes=(7, 5, 5), domain_lengths=(1.0, 0.8, 0.6))),
    "d1_2d": dict(nsd=2, sizes=(8, 8, 1), lengths=(1.0, 1.0, 1.0), deg=1, ngp=2, ctor=dict(domain_size=8, fem_basis_deg=1)),
}
KEYS = {"N": "gauss_pt_evaluation", "dx": "gauss_pt_evaluation_der_x", "dy": "gauss_pt_evaluation_der_y",
        "dz": "gauss_pt_evaluation_der_z", "surf": "gauss_pt_evaluation_surf"}


@pytest.mark.parametrize("tag", sorted(CASES))
def test_oracle_matches_the_reference(tag):
    c = CASES[tag]
    o = LagrangeOracle(c["nsd"], c["sizes"], c["lengths"], c["deg"], c["ngp"], dtype=torch.float32)
    u = torch.from_numpy(G[tag + ".u"])
    assert [o.ngp_1d, o.nbf_1d] == list(G[tag + ".meta"][:2])
    for k, name in KEYS.items():
        if tag + "." + k not in G.files:
            continue
        t = u[:, :, 0, :].contiguous() if k == "surf" else u
        out = getattr(o, name)(t)
        assert torch.equal(out, torch.from_numpy(G[tag + "." + k])), (tag, k)      # same convs, same fp32 stencils


def test_module_host_tables_for_degree_2_and_3():
    """Attributes the reference computes at construction (no GPU needed)."""
    from diffnet_b200 import DiffNet2DFEM, DiffNet3DFEM
    for tag, c in CASES.items():
        if c["deg"] == 1:
            continue
        m = (DiffNet3DFEM if c["nsd"] == 3 else DiffNet2DFEM)(None, **c["ctor"])
        assert [m.ngp_1d, m.nbf_1d, m.nelem] == list(G[tag + ".meta"])
        assert torch.equal(m.gpw, torch.from_numpy(G[tag + ".gpw"]))
        assert float((m.xgp - torch.from_numpy(G[tag + ".xgp"])).abs().max()) < 2e-6
        o = LagrangeOracle(c["nsd"], c["sizes"], c["lengths"], c["deg"], c["ngp"], dtype=torch.float32)
        for a, b in zip(m.N_gp, o.tables["N"]):
            assert torch.equal(a.data, b)
        for a, b in zip(m.dN_x_gp, o.tables["dx"]):
            assert torch.equal(a.data, b)
        with pytest.raises(NotImplementedError, match="Q1"):
            m.energy_loss(torch.zeros(1))
    with pytest.raises(AssertionError):
        DiffNet2DFEM(None, domain_size=8, fem_basis_deg=2)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", sorted(CASES))
def test_general_basis_cuda_ops(tag):
    """dn_fem_gp_eval_general_f32 (+ adjoint through autograd) against the reference outputs and the fp64 oracle."""
    from diffnet_b200 import DiffNet2DFEM, DiffNet3DFEM
    c = CASES[tag]
    dev = "cuda:0"
    m = (DiffNet3DFEM if c["nsd"] == 3 else DiffNet2DFEM)(None, **c["ctor"])
    o = LagrangeOracle(c["nsd"], c["sizes"], c["lengths"], c["deg"], c["ngp"], dtype=torch.float64)
    u = torch.from_numpy(G[tag + ".u"])
    gen = torch.Generator().manual_seed(5)
    for k, name in KEYS.items():
        if tag + "." + k not in G.files:
            continue
        t = u[:, :, 0, :].contiguous() if k == "surf" else u
        td = t.to(dev).requires_grad_(True)
        out = getattr(m, name)(td)
        ref = torch.from_numpy(G[tag + "." + k])
        assert out.shape == ref.shape
        assert rel_l2(out.detach().cpu(), ref) <= 1e-6, (tag, k)
        t64 = t.double().requires_grad_(True)
        o64 = getattr(o, name)(t64)
        cot = torch.randn(ref.shape, generator=gen)
        (out * cot.to(dev)).sum().backward()
        (o64 * cot.double()).sum().backward()
        assert rel_l2(td.grad.cpu(), t64.grad) <= 1e-6, (tag, k)
        assert float((td.grad.cpu().double() - t64.grad).abs().max() / t64.grad.abs().max()) <= 3e-6
