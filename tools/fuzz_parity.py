"""Randomised parity sweep on the GPU: random mesh sizes, batch sizes, views (channel slices,
broadcast fields), Dirichlet sets and options; streaming kernels vs the fp64 oracle (small
sizes) and vs the general kernels (all sizes).  python tools/fuzz_parity.py [seconds] [seed]"""
import os, sys, time, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from diffnet_b200 import DiffNet2DFEM, DiffNet3DFEM
from oracle import losses as OL
from oracle.fem import Q1Oracle

DEV = "cuda:0"
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
os.environ["DN_POISON_OUTPUTS"] = "1"


def rel(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float((a - b).norm() / max(float(b.norm()), 1e-300))


def case(i):
    nsd = rng.choice([2, 2, 3])
    if nsd == 2:
        nx = rng.choice([8, 12, 36, 64, 100, 132, 256, 260, 516, 1024, 2048, 2052]); ny = rng.randint(2, 70)
        sizes, dims = (nx, ny, 1), (ny, nx)
    else:
        nx = rng.choice([8, 12, 20, 36, 64, 72, 132, 256, 260, 516]); ny = rng.randint(2, 40); nz = rng.randint(2, 24)
        sizes, dims = (nx, ny, nz), (nz, ny, nx)
    B = rng.choice([1, 1, 2, 3, 5, 17])
    lengths = (1.0, rng.choice([1.0, 0.7]), rng.choice([1.0, 0.4]))
    cls = DiffNet2DFEM if nsd == 2 else DiffNet3DFEM
    ngp = rng.choice([2, 2, 2, 3])
    fem = cls(None, domain_sizes=sizes, domain_lengths=lengths, domain_size=nx, ngp_1d=ngp)
    g = torch.Generator().manual_seed(rng.randint(0, 1 << 30))
    shape = (B, 1) + dims
    u = torch.randn(shape, generator=g)
    # views: fields as channel slices of one tensor, or broadcast over the batch
    pack = torch.randn((B, 4) + dims, generator=g)
    nu = torch.exp(0.4 * pack[:, 0:1]) if rng.random() < 0.5 else torch.exp(0.4 * torch.randn((1, 1) + dims, generator=g))
    f = pack[:, 1:2] if rng.random() < 0.5 else torch.randn(shape, generator=g)
    m1 = (pack[:, 2:3] > 0.8).float(); m2 = (pack[:, 3:4] > 1.0).float()
    m3 = torch.zeros(shape); m3[..., 0] = 1
    nm = rng.choice([0, 1, 2, 3])
    dirichlet = [(m1, 1.0), (m2, 0.0), (m3, 0.25)][:nm]
    kw = {}
    if rng.random() < 0.8: kw["nu"] = nu
    if rng.random() < 0.7: kw["f"] = f
    elif rng.random() < 0.7:      # forcing at the Gauss points: assembled load vector (streaming) vs f_gp read (general)
        kw["f_gp"] = torch.randn((rng.choice([1, B]), ngp ** nsd) + tuple(d - 1 for d in dims), generator=g)
    if nm: kw["dirichlet"] = dirichlet
    if "nu" in kw and "f_gp" not in kw and rng.random() < 0.2: kw["nu_zero_mask"] = m3     # f_gp + nu mask: not offered
    if rng.random() < 0.3: kw["c_k"] = 0.5
    if rng.random() < 0.2: kw["scale"] = 0.37
    if rng.random() < 0.2: kw["reduction"] = "sum"
    dev = lambda t: t.to(DEV)
    if rng.random() < 0.15:
        return residual_case(i, fem, nsd, sizes, B, u, kw, dev)
    kwd = {k: (dev(v) if torch.is_tensor(v) else ([(dev(m), val) for m, val in v] if k == "dirichlet" else v)) for k, v in kw.items()}
    ud = dev(u)
    ls, gs = fem.energy_loss_and_grad(ud, **kwd)
    key = "DN_2D_PATH" if nsd == 2 else "DN_3D_PATH"
    os.environ[key] = "warp" if nsd == 2 else "tile"
    try:
        lg, gg = fem.energy_loss_and_grad(ud, **kwd)
    finally:
        os.environ.pop(key)
    torch.cuda.synchronize()
    desc = f"case {i}: nsd={nsd} sizes={sizes} B={B} ngp={ngp} opts={sorted(kw)} nm={nm}"
    assert torch.isfinite(gs).all() and torch.isfinite(ls), desc + " (non-finite / unwritten output)"
    el, eg = abs(float(ls) - float(lg)) / max(abs(float(lg)), 1e-30), rel(gs, gg)
    assert el < 5e-6 and eg < 5e-6, f"{desc}: streaming vs general loss {el:.2e} grad {eg:.2e}"
    ndof = B * dims[0] * dims[1] * (dims[2] if nsd == 3 else 1)
    if ndof <= 60000:
        o = Q1Oracle(nsd=nsd, domain_sizes=sizes, domain_lengths=lengths, domain_size=nx, ngp_1d=ngp, dtype=torch.float64)
        u64 = u.double().requires_grad_(True)
        kw64 = {k: (v.double() if torch.is_tensor(v) else ([(m.double(), val) for m, val in v] if k == "dirichlet" else v)) for k, v in kw.items()}
        lref = OL.energy_loss(o, u64, **kw64)
        (gref,) = torch.autograd.grad(lref, u64)
        el2 = abs(float(ls) - float(lref)) / max(abs(float(lref)), 1e-30)
        eg2 = rel(gs.reshape(gref.shape), gref) if float(gref.norm()) > 0 else float(gs.abs().max())
        assert el2 < 1e-5 and eg2 < 1e-4, f"{desc}: vs oracle loss {el2:.2e} grad {eg2:.2e}"
        for m, _ in kw.get("dirichlet", ()):
            assert float((gs.reshape(gref.shape).cpu() * (m > 0.5)).abs().max()) == 0.0, desc + " (gradient on a Dirichlet node)"
    return desc


def residual_case(i, fem, nsd, sizes, B, u, kw, dev):
    """sum(R^2) and its gradient (a second operator pass with mask_input = 0): streaming vs general."""
    rkw = {k: v for k, v in kw.items() if k in ("nu", "f", "dirichlet")}
    rkwd = {k: (dev(v) if torch.is_tensor(v) else [(dev(m), val) for m, val in v]) for k, v in rkw.items()}
    jac = (0.5 * fem.h) ** nsd
    out = []
    key = "DN_2D_PATH" if nsd == 2 else "DN_3D_PATH"
    for path in (None, "warp" if nsd == 2 else "tile"):
        if path:
            os.environ[key] = path
        try:
            ud = dev(u).clone().requires_grad_(True)
            loss = fem.residual_loss(ud, jac=jac, **rkwd)
            loss.backward()
            out.append((loss.detach(), ud.grad.detach()))
        finally:
            os.environ.pop(key, None)
    torch.cuda.synchronize()
    desc = f"case {i} (residual): nsd={nsd} sizes={sizes} B={B} opts={sorted(rkw)}"
    (l0, g0), (l1, g1) = out
    assert torch.isfinite(g0).all(), desc
    el = abs(float(l0) - float(l1)) / max(abs(float(l1)), 1e-30)
    eg = rel(g0, g1) if float(g1.norm()) > 0 else float(g0.abs().max())
    assert el < 1e-5 and eg < 1e-5, f"{desc}: loss {el:.2e} grad {eg:.2e}"
    return desc


t0, n = time.time(), 0
while time.time() - t0 < budget:
    case(n)
    n += 1
print(f"fuzz: {n} cases ok in {time.time() - t0:.0f} s")
