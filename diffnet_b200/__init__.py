"""diffnet_b200 -- the DiffNet FEM-loss hot path as hand-written sm_100a CUDA behind the
reference's own module API (DiffNetFEM / DiffNet2DFEM / DiffNet3DFEM, PDE).

Importing the package does not load CUDA; the first op call loads
``diffnet_b200/lib/libdiffnet_fem.so`` and raises if it is missing (no fallback).
"""
from .base import PDE
from .fem import DiffNet2DFEM, DiffNet3DFEM, DiffNetFEM
from .ops import (Geometry, PreparedEnergy, fem_energy, fem_energy_and_grad, fem_residual, gp_eval)

__all__ = ["PDE", "DiffNetFEM", "DiffNet2DFEM", "DiffNet3DFEM", "Geometry", "PreparedEnergy", "fem_energy",
           "fem_energy_and_grad", "fem_residual", "gp_eval"]
__version__ = "0.1.0"
