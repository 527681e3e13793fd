"""Generate tests/golden/highorder.npz from the REAL reference for fem_basis_deg 2 / 3 and the surface stencils:

    python tests/golden/make_golden_highorder.py        # build container (/root/reference mounted)

Upstream's degree 2 / 3 lambdas say ``dtype=np.float``, removed in numpy 1.24: the alias is restored for this
script (``np.float = float``) -- an environment shim like the Lightning stub, the reference files are untouched.
"""
import os
import sys

import numpy as np
import torch

np.float = float            # noqa: the alias numpy < 1.24 had; DiffNetFEM.py:75,80,85,113,119,126 use it
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.refload import load_reference  # noqa: E402

ref = load_reference()
torch.manual_seed(20261019)
out = {}
for tag, cls, kw, shape in (
        ("d2_2d", ref.DiffNet2DFEM, dict(domain_size=9, fem_basis_deg=2, domain_sizes=(9, 7, 1), domain_lengths=(1.2, 0.9, 1.0)), (2, 1, 7, 9)),
        ("d3_2d", ref.DiffNet2DFEM, dict(domain_size=10, fem_basis_deg=3, ngp_1d=4), (1, 1, 10, 10)),
        ("d2_3d", ref.DiffNet3DFEM, dict(domain_size=7, fem_basis_deg=2, nsd=3, domain_sizes=(7, 5, 5), domain_lengths=(1.0, 0.8, 0.6)), (1, 1, 5, 5, 7)),
        ("d1_2d", ref.DiffNet2DFEM, dict(domain_size=8, fem_basis_deg=1), (2, 1, 8, 8))):
    m = cls(None, **kw)
    u = torch.randn(*shape)
    out[tag + ".u"] = u.numpy()
    out[tag + ".N"] = m.gauss_pt_evaluation(u).detach().numpy()
    out[tag + ".dx"] = m.gauss_pt_evaluation_der_x(u).detach().numpy()
    out[tag + ".dy"] = m.gauss_pt_evaluation_der_y(u).detach().numpy()
    if m.nsd == 3:
        out[tag + ".dz"] = m.gauss_pt_evaluation_der_z(u).detach().numpy()
    else:
        line = u[:, :, 0, :].contiguous()
        out[tag + ".surf"] = m.gauss_pt_evaluation_surf(line).detach().numpy()
    out[tag + ".gpw"] = m.gpw.numpy()
    out[tag + ".xgp"] = m.xgp.numpy()
    out[tag + ".meta"] = np.array([m.ngp_1d, m.nbf_1d, m.nelem])
np.savez_compressed(os.path.join(HERE, "highorder.npz"), **out)
print({k: v.shape for k, v in out.items()})
