"""A minimal trainer for ``PDE`` modules: the slice of ``pytorch_lightning.Trainer.fit`` the
reference scripts use (automatic optimisation: training_step -> backward -> optimizer.step;
``strategy='ddp'`` = one process per GPU with gradient all-reduce, IBN/poisson-3d/parametric/
IBN_3D.py:193-205).  PyTorch-Lightning is not in this image; when it is installed the modules
work with the real Trainer unchanged (they subclass its LightningModule).

Data-parallel path (parametric case): the batch is sharded across ranks, the FEM
loss kernel runs unmodified on each rank's shard (local mean), and ``DistributedDataParallel``
averages the network gradients with bucketed NCCL all-reduces over NVLink that overlap the rest
of the backward pass -- the same arithmetic as Lightning's DDP strategy.  No collective touches
the FEM path itself.
"""
from __future__ import annotations

import os
import time
from typing import Iterable, Optional

import torch
import torch.distributed as dist


def init_distributed(backend: Optional[str] = None) -> tuple[int, int, int]:
    """(rank, world, local_rank) from the torchrun environment; initialises the default process
    group when WORLD_SIZE > 1 (NCCL on CUDA, gloo on CPU)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def shard_batch(batch, rank: int, world: int):
    """This rank's contiguous slice of a global batch (what DistributedSampler + DataLoader give
    Lightning's DDP strategy)."""
    def cut(t):
        n = t.shape[0]
        if n % world:
            raise ValueError(f"global batch {n} is not divisible by world size {world}")
        per = n // world
        return t[rank * per:(rank + 1) * per]
    return tuple(cut(t) for t in batch)


class Trainer:
    """fit(module, batches): automatic optimisation over an iterable of (inputs, forcing) batches
    that are already this rank's shard.  With ``world > 1`` the module's network is wrapped in
    DistributedDataParallel (gradient averaging == the global-mean loss for equal shards)."""

    def __init__(self, max_steps: int = 100, device: Optional[torch.device] = None, ddp: Optional[bool] = None,
                 bucket_cap_mb: int = 25, log_every: int = 0):
        self.max_steps, self.log_every, self.bucket_cap_mb = max_steps, log_every, bucket_cap_mb
        self.rank, self.world, self.local = init_distributed() if ddp is not False else (0, 1, 0)
        if device is None:
            device = torch.device("cuda", self.local) if torch.cuda.is_available() else torch.device("cpu")
        self.device = device
        self.ddp = (self.world > 1) if ddp is None else (ddp and self.world > 1)
        self.losses = []

    def _wrap(self, module):
        module.to(self.device)
        net = module.network
        if self.ddp and isinstance(net, torch.nn.Module) and any(p.requires_grad for p in net.parameters()):
            ids = [self.device.index] if self.device.type == "cuda" else None
            module.network = torch.nn.parallel.DistributedDataParallel(
                net, device_ids=ids, bucket_cap_mb=self.bucket_cap_mb, gradient_as_bucket_view=True)
        return module

    def fit(self, module, batches: Iterable):
        module = self._wrap(module)
        module.train()
        opts, _ = module.configure_optimizers()
        opt = opts[0]
        step = 0
        t0 = time.perf_counter()
        for batch in batches:
            if step >= self.max_steps:
                break
            batch = tuple(t.to(self.device, non_blocking=True) for t in batch)
            opt.zero_grad(set_to_none=True)
            loss = module.training_step(batch, step)
            loss.backward()
            opt.step()
            self.losses.append(loss.detach())
            step += 1
            if self.log_every and self.rank == 0 and step % self.log_every == 0:
                print(f"[trainer] step {step} loss {float(loss):.6g} ({time.perf_counter() - t0:.1f} s)", flush=True)
        return module
