"""Generate tests/golden/producers.npz from the REAL reference's dataset code (adityabalu/DiffNet):

    python tests/golden/make_golden_producers.py        # in the build container (/root/reference mounted)

* ``DiffNet.gen_input_calc.generate_diffusivity_tensor`` / ``calculate_omega_based_on_eta`` (2-D and 3-D KL fields)
  wrapped exactly like ``DiffNet/datasets/parametric/klsum.py:21-33`` builds [nu, bc1, bc2];
* ``DiffNet.datasets.parametric.images.ImageIMBack`` on two small PNGs written to a temporary directory;
* ``DiffNet.datasets.single_instances.voxels.VoxelIMBackRAW`` on a small raw voxel file + VoxelConfig.txt.
Inputs (coefficients, image bytes, raw bytes) are stored beside the outputs so that the GPU producers
(diffnet_b200/csrc/producers.cu) can be checked on a box that has no reference tree.
"""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.refload import load_reference  # noqa: E402

load_reference()                                             # puts the reference on sys.path (Lightning stubbed)
from DiffNet.gen_input_calc import calculate_omega_based_on_eta, generate_diffusivity_tensor  # noqa: E402
from DiffNet.datasets.parametric.images import ImageIMBack  # noqa: E402
from DiffNet.datasets.single_instances.voxels import VoxelIMBackRAW  # noqa: E402

rng = np.random.default_rng(20261018)
out = {}

# ---- KL fields, 2-D: the three channels KLSumStochastic stores (klsum.py:21-33), as FloatTensor does (fp32)
coeffs = rng.uniform(-3.0, 3.0, size=(3, 6))
N = 16
samples = []
for c in coeffs:
    domain = generate_diffusivity_tensor(c, output_size=N, n_sum_nu=6).squeeze()
    bc1 = np.zeros_like(domain); bc1[:, 0] = 1
    bc2 = np.zeros_like(domain); bc2[:, -1] = 1
    samples.append(torch.FloatTensor(np.array([domain, bc1, bc2])).numpy())
out["kl.coeffs"] = coeffs
out["kl.omega"] = calculate_omega_based_on_eta(0.5)
out["kl2d.inputs"] = np.stack(samples)
# ---- KL field, 3-D
N3 = 8
out["kl3d.nu"] = np.stack([torch.FloatTensor(generate_diffusivity_tensor(c, output_size=N3, nsd=3)).numpy() for c in coeffs[:2]])

# ---- ImageIMBack on two PNGs
import PIL.Image  # noqa: E402
imgs = (rng.uniform(0, 1, size=(2, 10, 12)) > 0.6).astype(np.uint8) * rng.integers(1, 255, size=(2, 10, 12), dtype=np.uint8)
with tempfile.TemporaryDirectory() as d:
    for i, im in enumerate(imgs):
        PIL.Image.fromarray(im, mode="L").save(os.path.join(d, f"img{i}.png"))
    ds = ImageIMBack(d, domain_size=12)
    items = [ds[i] for i in range(len(ds))]
out["img.bytes"] = imgs
out["img.inputs"] = np.stack([it[0].numpy() for it in items])
out["img.forcing"] = np.stack([it[1].numpy() for it in items])

# ---- VoxelIMBackRAW on a small raw file (the reference places the block at offset 32)
d0, d1, d2 = 5, 4, 3
raw = rng.integers(0, 255, size=d0 * d1 * d2, dtype=np.uint8)
with tempfile.TemporaryDirectory() as d:
    base = os.path.join(d, "bunny_")
    raw.tofile(base + "inouts.raw")
    with open(base + "VoxelConfig.txt", "w") as fh:
        fh.write("header\n0 0 0\n1 1 1\n%d %d %d\n0.1 0.1 0.1\n10\n5\n" % (d0, d1, d2))
    ds = VoxelIMBackRAW(base, domain_size=40)
    inp, frc = ds[0]
out["vox.raw"] = raw
out["vox.num_div"] = np.array([d0, d1, d2])
out["vox.inputs"] = inp.numpy()[None]
out["vox.forcing"] = frc.numpy()[None]

np.savez_compressed(os.path.join(HERE, "producers.npz"), **out)
print({k: (v.shape, str(v.dtype)) for k, v in out.items()})
