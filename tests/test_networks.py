"""The UNet stays ordinary PyTorch (north_star); this only checks that the restated module is the
reference's architecture: parameter count, shapes, and -- where /root/reference is mounted --
identical outputs for identical weights (DiffNet/networks/unets.py:13-81)."""
import pytest
import torch

from diffnet_b200.networks import UNet
from oracle.refload import reference_available


def test_unet_shapes_and_parameter_count():
    net = UNet(2, 1)
    assert sum(p.numel() for p in net.parameters()) == 4_163_585
    assert net(torch.randn(2, 2, 64, 64)).shape == (2, 1, 64, 64)
    net3 = UNet(1, 1, nd=3)
    assert abs(sum(p.numel() for p in net3.parameters()) - 4_163_585) < 2_000
    assert net3(torch.randn(1, 1, 64, 64, 64)).shape == (1, 1, 64, 64, 64)


@pytest.mark.skipif(not reference_available(), reason="/root/reference not mounted")
def test_unet_equals_reference_function():
    import sys
    from oracle.refload import load_reference
    load_reference()                                   # installs the pytorch_lightning stub
    sys.path.insert(0, "/root/reference")
    from DiffNet.networks.unets import UNet as RefUNet
    torch.manual_seed(0)
    ref, mine = RefUNet(2, 1), UNet(2, 1)
    with torch.no_grad():
        for a, b in zip(mine.parameters(), ref.parameters()):
            a.copy_(b)
    ref.eval(); mine.eval()
    x = torch.randn(2, 2, 64, 64)
    assert torch.equal(mine(x), ref(x))
