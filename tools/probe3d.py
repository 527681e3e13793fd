import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffnet_b200 import DiffNet3DFEM
B, D, H, W = [int(v) for v in sys.argv[1:5]]
fem = DiffNet3DFEM(None, domain_sizes=(W, H, D), domain_size=W)
g = torch.Generator(device="cuda").manual_seed(0)
u = torch.randn(B, 1, D, H, W, device="cuda", generator=g)
nu = torch.rand(B, 1, D, H, W, device="cuda", generator=g) + 0.5
try:
    l, gr = fem.energy_loss_and_grad(u, nu=nu)
    torch.cuda.synchronize()
    os.environ["DN_3D_PATH"] = "tile"
    l2, g2 = fem.energy_loss_and_grad(u, nu=nu)
    torch.cuda.synchronize()
    print("ok", float(l), float(l2), float((gr - g2).abs().max() / g2.abs().max()))
except Exception as e:
    print("FAIL", str(e).splitlines()[0])
