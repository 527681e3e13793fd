"""Device-side input producers (csrc/producers.cu, diffnet_b200/datasets.py) against tensors produced by the
REAL reference's dataset code (tests/golden/producers.npz, written by tests/golden/make_golden_producers.py)."""
import os

import numpy as np
import pytest
import torch

from conftest import ROOT
from diffnet_b200 import DiffNet2DFEM, datasets as D
from diffnet_b200.synthetic import _star_raster_cpu, box_params, poisson3d_parametric_batch, star_params

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
G = np.load(os.path.join(ROOT, "tests", "golden", "producers.npz"))


def ulp_close(a, b, ulps=1.0):
    a, b = a.double(), b.double()
    return bool(((a - b).abs() <= ulps * 1.1920929e-07 * b.abs()).all())


def test_kl_inputs_2d_match_the_reference_dataset():
    ref = torch.from_numpy(G["kl2d.inputs"])
    inputs, forcing = D.kl_inputs(G["kl.coeffs"], 16, 2, 0.5, G["kl.omega"], DEV)
    out = inputs.cpu()
    assert out.shape == ref.shape and forcing.shape == (3, 1, 16, 16) and float(forcing.abs().max()) == 0.0
    assert torch.equal(out[:, 1:], ref[:, 1:])                          # bc1 / bc2 bit for bit
    assert ulp_close(out[:, 0], ref[:, 0], 1.0)                         # nu: fp64 sum + exp, rounded to fp32
    assert float((out[:, 0] == ref[:, 0]).float().mean()) > 0.99
    # roots solved here instead of the reference's table: same field to 2 ulp
    inputs2, _ = D.kl_inputs(G["kl.coeffs"], 16, 2, 0.5, None, DEV)
    assert ulp_close(inputs2[:, 0].cpu(), ref[:, 0], 2.0)


def test_kl_field_3d_matches_the_reference_generator():
    ref = torch.from_numpy(G["kl3d.nu"])
    nu, _ = D.kl_inputs(G["kl.coeffs"][:2], 8, 3, 0.5, G["kl.omega"], DEV)
    assert nu.shape == ref.shape
    assert ulp_close(nu.cpu(), ref, 1.0)


def test_image_and_voxel_inputs_match_the_reference_datasets():
    inputs, forcing = D.image_inputs(G["img.bytes"], DEV)
    assert torch.equal(inputs.cpu(), torch.from_numpy(G["img.inputs"]))
    assert torch.equal(forcing.cpu(), torch.from_numpy(G["img.forcing"]))
    vin, vf = D.voxel_inputs(torch.from_numpy(G["vox.raw"]), G["vox.num_div"], 40, 32, DEV)
    assert torch.equal(vin.cpu(), torch.from_numpy(G["vox.inputs"]))
    assert torch.equal(vf.cpu(), torch.from_numpy(G["vox.forcing"]))
    with pytest.raises(Exception, match="does not fit"):
        D.voxel_inputs(torch.from_numpy(G["vox.raw"]), G["vox.num_div"], 32, 32, DEV)


def test_dataset_classes_feed_the_fused_loss(tmp_path):
    """KLSumStochastic with the reference's constructor arguments: items, batches, and a loss step on them."""
    np.save(tmp_path / "sobol.npy", G["kl.coeffs"])
    ds = D.KLSumStochastic(str(tmp_path / "sobol.npy"), domain_size=16, kl_terms=6, device=DEV, omega=G["kl.omega"])
    assert len(ds) == 3
    inp0, f0 = ds[1]
    assert inp0.shape == (3, 16, 16) and f0.shape == (1, 16, 16) and inp0.is_cuda
    assert ulp_close(inp0.cpu(), torch.from_numpy(G["kl2d.inputs"][1]), 1.0)
    inputs, forcing = ds.batch([0, 1, 2])
    fem = DiffNet2DFEM(None, domain_size=16)
    u = torch.rand(3, 1, 16, 16, device=DEV, requires_grad=True)
    loss = fem.energy_loss(u, nu=inputs[:, 0:1], f=forcing, dirichlet=[(inputs[:, 1:2], 1.0), (inputs[:, 2:3], 0.0)])
    loss.backward()
    assert torch.isfinite(loss) and float(u.grad[:, :, :, 0].abs().max()) == 0.0     # Dirichlet column: no gradient


def test_synthetic_geometry_kernels_match_the_host_rasterisers():
    P, _ = star_params(4, seed=3)
    inputs, forcing = D.star_inputs(P, 128, DEV)
    obj = _star_raster_cpu(P, 128)
    mism = float((inputs[:, 1:2].cpu() != obj).float().mean())
    assert mism < 5e-4                                                   # borderline pixels: last-ulp trig differences
    assert torch.equal(inputs[:, 0], 1.0 - inputs[:, 1]) and float(forcing.abs().max()) == 0.0
    edge = inputs[:, 2].cpu()
    assert float(edge[:, 0].min()) == 1.0 and float(edge[:, :, -1].min()) == 1.0 and float(edge[:, 1:-1, 1:-1].max()) == 0.0
    u, src, sink, f = poisson3d_parametric_batch(3, 32, DEV, seed=5)
    uc, srcc, sinkc, fc = poisson3d_parametric_batch(3, 32, "cpu", seed=5)
    assert torch.equal(src.cpu(), srcc) and torch.equal(sink.cpu(), sinkc) and torch.equal(f.cpu(), fc)
    assert torch.equal(u.cpu(), uc)
