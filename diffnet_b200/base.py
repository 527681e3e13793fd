"""PDE base module -- same constructor kwargs, attributes and hooks as the reference's
``DiffNet/base.py:6-55`` so that user subclasses (``class Poisson(DiffNet2DFEM)`` with a
``loss()``) keep working.  Uses real PyTorch-Lightning when it is importable; otherwise a
minimal stand-in (``nn.Module`` + ``log``) that ``diffnet_b200.trainer.Trainer`` can drive.
"""
from __future__ import annotations

import torch

try:                                            # pragma: no cover - not installed in this image
    from pytorch_lightning.core import LightningModule as _Base
    HAVE_LIGHTNING = True
except Exception:                               # noqa: BLE001
    HAVE_LIGHTNING = False

    class _Base(torch.nn.Module):
        """The slice of LightningModule the reference uses: ``log`` and the step hooks."""

        def __init__(self):
            super().__init__()
            self.logged = {}

        def log(self, name, value, *args, **kwargs):
            self.logged[name] = value


class PDE(_Base):
    """kwargs -> geometry, exactly as DiffNet/base.py:11-32."""

    def __init__(self, network, **kwargs):
        super().__init__()
        self.kwargs = kwargs
        self.network = network
        self.nsd = kwargs.get("nsd", 2)
        self.batch_size = kwargs.get("batch_size", 64)
        self.n_workers = kwargs.get("n_workers", 1)
        self.learning_rate = kwargs.get("learning_rate", 3e-4)

        self.domain_length = kwargs.get("domain_length", 1.0)
        self.domain_size = kwargs.get("domain_size", 64)
        L, N = self.domain_length, self.domain_size
        self.domain_lengths_nd = kwargs.get("domain_lengths", (L, L, L))
        self.domain_sizes_nd = kwargs.get("domain_sizes", (N, N, N))
        if self.nsd >= 2:
            self.domain_lengthX, self.domain_lengthY = self.domain_lengths_nd[0], self.domain_lengths_nd[1]
            self.domain_sizeX, self.domain_sizeY = self.domain_sizes_nd[0], self.domain_sizes_nd[1]
        if self.nsd >= 3:
            self.domain_lengthZ = self.domain_lengths_nd[2]
            self.domain_sizeZ = self.domain_sizes_nd[2]

    def loss(self, u, inputs_tensor, forcing_tensor):
        raise NotImplementedError

    def forward(self, batch):
        inputs_tensor, forcing_tensor = batch
        u = self.network(inputs_tensor)
        return u, inputs_tensor, forcing_tensor

    def training_step(self, batch, batch_idx):
        u, inputs_tensor, forcing_tensor = self.forward(batch)
        loss_val = self.loss(u, inputs_tensor, forcing_tensor).mean()
        # The reference logs loss_val.item() here (base.py:45-46), a device sync per step.
        # The detached tensor is logged instead; Lightning and our Trainer read it lazily.
        self.log("PDE_loss", loss_val.detach())
        self.log("loss", loss_val.detach())
        return loss_val

    def configure_optimizers(self):
        lr = self.learning_rate
        opts = [torch.optim.Adam(self.network.parameters(), lr=lr)]
        return opts, []
