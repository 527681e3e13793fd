"""Oracle (TEST INFRASTRUCTURE): import the real reference module when it is mounted.

``/root/reference`` exists only in the build container (never on the GPU box).
``DiffNet/base.py:3`` imports ``pytorch_lightning.core.LightningModule``; Lightning is
not installed in this image, so a 6-line stand-in is injected first.
Used by ``tests/test_oracle_vs_reference.py`` and ``tests/golden/make_golden.py``.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("DIFFNET_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "DiffNet", "DiffNetFEM.py"))


def load_reference():
    """Returns the reference's ``DiffNet.DiffNetFEM`` module (unmodified)."""
    if not reference_available():
        raise FileNotFoundError(REFERENCE_ROOT)
    import torch

    if "pytorch_lightning" not in sys.modules:
        class _LightningModule(torch.nn.Module):
            def log(self, *a, **k):
                pass
        pl = types.ModuleType("pytorch_lightning")
        core = types.ModuleType("pytorch_lightning.core")
        core.LightningModule = pl.LightningModule = _LightningModule
        pl.core = core
        sys.modules["pytorch_lightning"] = pl
        sys.modules["pytorch_lightning.core"] = core
    if REFERENCE_ROOT not in sys.path:
        sys.path.append(REFERENCE_ROOT)
    import DiffNet.DiffNetFEM as ref          # noqa: E402
    return ref
