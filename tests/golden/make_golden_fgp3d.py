"""Generate tests/golden/ref_3d_fgp.npz from the REAL reference (adityabalu/DiffNet): the 3-D forcing-at-Gauss-points
form of examples/poisson/mms/e8_3d_poisson_mms.py (u = where(bc, u_bc, u); 0.5 nu |grad u|^2 - f_gp u_gp integrated
with the reference's own gauss_pt_evaluation* and tables), fp32 and fp64, like make_golden.py.

    python tests/golden/make_golden_fgp3d.py          (build container: /root/reference mounted)

A separate script so that the other fixtures (one shared RNG stream in make_golden.py) stay byte-identical.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import losses as L            # noqa: E402
from oracle.refload import load_reference  # noqa: E402

ref = load_reference()
torch.manual_seed(20261019)
torch.set_num_threads(1)


def grad_of(fn, u):
    u = u.clone().requires_grad_(True)
    loss = fn(u)
    (g,) = torch.autograd.grad(loss, u)
    return loss.detach(), g


def main():
    X, Y, Z, B = 8, 6, 5, 2
    mk = lambda: ref.DiffNet3DFEM(None, nsd=3, domain_sizes=(X, Y, Z), domain_lengths=(1.0, 0.8, 0.5), domain_size=X)
    u = torch.randn(B, 1, Z, Y, X)
    nu = torch.exp(0.5 * torch.randn(B, 1, Z, Y, X))
    bc = torch.zeros(1, 1, Z, Y, X)
    bc[:, :, 0] = 1; bc[:, :, -1] = 1; bc[:, :, :, 0] = 1; bc[:, :, :, -1] = 1; bc[..., 0] = 1; bc[..., -1] = 1
    u_bc = torch.randn(1, 1, Z, Y, X)
    f_gp = torch.randn(1, 8, Z - 1, Y - 1, X - 1)
    body = lambda fm, v, nu_, fg_, ub_: L.energy_loss(fm, v, nu=nu_, f_gp=fg_, dirichlet=[(bc.to(v.dtype), ub_)], c_k=0.5)
    out = dict(u=u.numpy(), nu=nu.numpy(), bc=bc.numpy(), u_bc=u_bc.numpy(), f_gp=f_gp.numpy(),
               sizes=np.array([X, Y, Z]), lengths=np.array([1.0, 0.8, 0.5]))
    l32, g32 = grad_of(lambda v: body(mk(), v, nu, f_gp, u_bc), u)
    l64, g64 = grad_of(lambda v: body(mk().double(), v, nu.double(), f_gp.double(), u_bc.double()), u.double())
    out.update({"E3fgp.loss": l32.numpy(), "E3fgp.grad": g32.numpy(), "E3fgp.loss64": l64.numpy(), "E3fgp.grad64": g64.numpy()})
    path = os.path.join(HERE, "ref_3d_fgp.npz")
    np.savez_compressed(path, **out)
    print(f"ref_3d_fgp: {os.path.getsize(path)} bytes, keys={len(out)}")


if __name__ == "__main__":
    main()
