"""The reference's Poisson ``loss()`` bodies, each as one fused call.

Every class is what a user of the reference writes in their script (``class Poisson(DiffNet2DFEM)``
with ``loss``/``forward``/``configure_optimizers``); the bodies cite the script they replace.
They are the callers of the hot path that the train-step benchmark and the multi-GPU paths drive.
"""
from __future__ import annotations

import torch

from .fem import DiffNet2DFEM, DiffNet3DFEM


class PoissonParametric2D(DiffNet2DFEM):
    """Parametric 2-D Poisson, KL diffusivity (examples/poisson/parametric/2_klsum_fem.py:33-62,
    12_klsum.py:53-78): inputs (B,3,H,W) = [nu, bc1, bc2], u = network(inputs), energy E1."""

    def loss(self, u, inputs_tensor, forcing_tensor):
        nu, bc1, bc2 = inputs_tensor[:, 0:1], inputs_tensor[:, 1:2], inputs_tensor[:, 2:3]
        return self.energy_loss(u, nu=nu, f=forcing_tensor, dirichlet=[(bc1, 1.0), (bc2, 0.0)])


class PoissonSingleInstance2D(DiffNet2DFEM):
    """Single-instance 2-D Poisson (examples/poisson/single_instance/0_base.py:31-56): energy E2,
    scale 0.5 (h/2)^2."""

    def loss(self, u, inputs_tensor, forcing_tensor):
        nu, bc1, bc2 = inputs_tensor[:, 0:1], inputs_tensor[:, 1:2], inputs_tensor[:, 2:3]
        return self.energy_loss(u, nu=nu, f=forcing_tensor, dirichlet=[(bc1, 1.0), (bc2, 0.0)],
                                scale=0.5 * (0.5 * self.h) ** 2)


class PoissonIBN2D(DiffNet2DFEM):
    """Immersed-background 2-D Poisson on images (IBN/poisson-2d/parametric/
    e1_complex_immersed_background.py:33-58): inputs [domain(nu), bc1 (object), bc2 (edges)];
    the network sees the first two channels (:60-63)."""

    def forward(self, batch):
        inputs_tensor, forcing_tensor = batch
        return self.network(inputs_tensor[:, 0:2]), inputs_tensor, forcing_tensor

    def loss(self, u, inputs_tensor, forcing_tensor):
        nu, bc1, bc2 = inputs_tensor[:, 0:1], inputs_tensor[:, 1:2], inputs_tensor[:, 2:3]
        return self.energy_loss(u, nu=nu, f=forcing_tensor, dirichlet=[(bc1, 1.0), (bc2, 0.0)])


class PoissonIBN3D(DiffNet3DFEM):
    """Parametric 3-D source/sink Poisson (IBN/poisson-3d/parametric/IBN_3D.py:105-160): batches are
    (source, sink, forcing), each (B,1,D,H,W); nu == 1; u = network(source).  The reference sets
    u = 1 on the source, removes the source from the sink (:119-121) and then sets u = 0 on what
    is left of the sink: for 0/1 masks that is "sink first, source last (wins)" in the ordered
    Dirichlet list."""

    def forward(self, batch):
        source_tensor, sink_tensor, forcing_tensor = batch
        return self.network(source_tensor), source_tensor, sink_tensor, forcing_tensor

    def loss(self, u, source_tensor, sink_tensor, forcing_tensor):
        return self.energy_loss(u, f=forcing_tensor, dirichlet=[(sink_tensor, 0.0), (source_tensor, 1.0)])

    def training_step(self, batch, batch_idx):
        u, source_tensor, sink_tensor, forcing_tensor = self.forward(batch)
        loss_val = self.loss(u, source_tensor, sink_tensor, forcing_tensor).mean()
        self.log("loss", loss_val.detach())
        return loss_val


class PoissonInObject3D(DiffNet3DFEM):
    """Non-parametric 3-D Poisson inside a voxelised object (IBN/poisson-3d/non-parametric/
    solve_in_object_3d.py:75-123): u IS the parameter (network = ParameterList([u])), bc1 = outside,
    nu = object mask, f = 500, energy E3 (c_k = 0.5), Adam(lr) on u."""

    def forward(self, batch):
        inputs_tensor, forcing_tensor = batch
        return self.network[0], inputs_tensor, forcing_tensor

    def loss(self, u, inputs_tensor, forcing_tensor):
        nu, bc1 = inputs_tensor[:, 0:1], inputs_tensor[:, 1:2]
        return self.energy_loss(u, nu=nu, f=forcing_tensor, dirichlet=[(bc1, 0.0)], c_k=0.5)

    def configure_optimizers(self):
        return [torch.optim.Adam(self.network.parameters(), lr=self.learning_rate)], []
