"""Device-side producers of the loss path's inputs behind the reference's dataset names.

The reference builds every input field on the HOST with numpy (``DiffNet/datasets/*``) and ships three fp32
channels per sample through a DataLoader and a host->device copy each step.  Here the same tensors are
written by CUDA kernels (``csrc/producers.cu``) straight into device memory:

  KLSumStochastic   DiffNet/datasets/parametric/klsum.py:10-46 over DiffNet/gen_input_calc.py:74-181
  ImageIMBack       DiffNet/datasets/parametric/images.py:9-49       (image decoding stays on the host: bytes in)
  VoxelIMBackRAW    DiffNet/datasets/single_instances/voxels.py:8-61 (the .raw bytes and VoxelConfig.txt in)

Each class keeps the reference's constructor arguments and ``__len__`` / ``__getitem__`` (so a DataLoader still
works: items are device tensors) and adds ``batch(indices)`` -> ``(inputs (B,3,...), forcing (B,1,...))``
generated on the device in one launch: what goes over PCIe per step is the sample parameters (48 B per KL
sample, 1 B per pixel for images), not 12-16 B per node.  There is no CPU path: ``device`` must be CUDA.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Sequence

import numpy as np
import torch

from . import _lib as L
from .synthetic import kl_omegas


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _need_cuda(device) -> torch.device:
    dev = torch.device(device)
    if dev.type != "cuda":
        raise L.DiffNetFEMError("diffnet_b200.datasets generates on the GPU: device must be CUDA (no CPU fallback)")
    return dev


def kl_inputs(coeffs, size: int, nsd: int = 2, eta: float = 0.5, omega: Sequence[float] | None = None,
              device="cuda", with_forcing: bool = True):
    """``coeffs`` (B, nterms) -> (inputs, forcing).  2-D: inputs (B,3,size,size) = [exp(KL sum), first column,
    last column]; 3-D: inputs (B,1,size,size,size) = the diffusivity (axes (y, x, z), gen_input_calc.py:124-130).
    ``omega``: the nterms roots for ``eta`` (default: solved here to 1e-16; pass the reference's table for
    bit-faithful values)."""
    dev = _need_cuda(device)
    c = torch.as_tensor(np.asarray(coeffs, dtype=np.float64) if not torch.is_tensor(coeffs) else coeffs)
    c = c.to(device=dev, dtype=torch.float64).contiguous()
    if c.dim() == 1:
        c = c[None]
    B, nterms = c.shape
    om = np.asarray(omega if omega is not None else kl_omegas(eta, nterms), dtype=np.float64)[:nterms]
    if om.shape[0] != nterms:
        raise ValueError(f"need {nterms} roots, got {om.shape[0]}")
    lib = L.lib()
    shape = (B, 3, size, size) if nsd == 2 else (B, 1, size, size, size)
    inputs = torch.empty(shape, device=dev, dtype=torch.float32)
    forcing = torch.empty((B, 1) + shape[2:], device=dev, dtype=torch.float32) if with_forcing else None
    tb = lib.dn_gen_kl_table_bytes(nterms, size)
    tables = torch.empty(max(tb, 8) // 8, device=dev, dtype=torch.float64)
    om_c = (C.c_double * nterms)(*om.tolist())
    with torch.cuda.device(dev):
        rc = lib.dn_gen_kl_inputs_f32(C.c_void_p(c.data_ptr()), B, nterms, om_c, float(eta), nsd, size,
                                      C.c_void_p(tables.data_ptr()), tb, C.c_void_p(inputs.data_ptr()),
                                      C.c_void_p(forcing.data_ptr()) if forcing is not None else None, _stream(dev))
    L.check(rc, "dn_gen_kl_inputs_f32")
    return inputs, forcing


def image_inputs(images, device="cuda", with_forcing: bool = True):
    """Greyscale images (B, H, W) uint8 (numpy or tensor) -> ([domain, bc1, bc2], forcing) like ImageIMBack."""
    dev = _need_cuda(device)
    im = torch.as_tensor(images)
    if im.dtype != torch.uint8:
        raise ValueError("images must be uint8 (PIL convert('L') bytes)")
    if im.dim() == 2:
        im = im[None]
    im = im.to(dev).contiguous()
    B, H, W = im.shape
    inputs = torch.empty((B, 3, H, W), device=dev, dtype=torch.float32)
    forcing = torch.empty((B, 1, H, W), device=dev, dtype=torch.float32) if with_forcing else None
    with torch.cuda.device(dev):
        rc = L.lib().dn_gen_image_inputs_f32(C.c_void_p(im.data_ptr()), B, H, W, C.c_void_p(inputs.data_ptr()),
                                             C.c_void_p(forcing.data_ptr()) if forcing is not None else None, _stream(dev))
    L.check(rc, "dn_gen_image_inputs_f32")
    return inputs, forcing


def voxel_inputs(raw, num_div, domain_size: int, offset: int = 32, device="cuda", with_forcing: bool = True):
    """The bytes of ``<name>inouts.raw`` + numDiv -> (inputs (1,3,N,N,N), forcing (1,1,N,N,N)) like VoxelIMBackRAW."""
    dev = _need_cuda(device)
    r = torch.as_tensor(np.frombuffer(raw, dtype=np.uint8).copy() if isinstance(raw, (bytes, bytearray)) else raw)
    if r.dtype != torch.uint8:
        raise ValueError("raw voxels must be uint8")
    d0, d1, d2 = (int(v) for v in num_div)
    if r.numel() != d0 * d1 * d2:
        raise ValueError(f"raw has {r.numel()} bytes, numDiv says {d0 * d1 * d2}")
    r = r.reshape(-1).to(dev).contiguous()
    N = int(domain_size)
    inputs = torch.empty((1, 3, N, N, N), device=dev, dtype=torch.float32)
    forcing = torch.empty((1, 1, N, N, N), device=dev, dtype=torch.float32) if with_forcing else None
    with torch.cuda.device(dev):
        rc = L.lib().dn_gen_voxel_inputs_f32(C.c_void_p(r.data_ptr()), d0, d1, d2, N, int(offset), C.c_void_p(inputs.data_ptr()),
                                             C.c_void_p(forcing.data_ptr()) if forcing is not None else None, _stream(dev))
    L.check(rc, "dn_gen_voxel_inputs_f32")
    return inputs, forcing


def star_inputs(params, size: int, device="cuda"):
    """Synthetic star-shaped silhouettes: params (B, 11) float32 = (cx, cy, r0, a[4], phase[4]) on [-1,1]^2."""
    dev = _need_cuda(device)
    p = torch.as_tensor(params, dtype=torch.float32).to(dev).contiguous()
    B = p.shape[0]
    inputs = torch.empty((B, 3, size, size), device=dev, dtype=torch.float32)
    forcing = torch.empty((B, 1, size, size), device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        rc = L.lib().dn_gen_star_inputs_f32(C.c_void_p(p.data_ptr()), B, size, C.c_void_p(inputs.data_ptr()),
                                            C.c_void_p(forcing.data_ptr()), _stream(dev))
    L.check(rc, "dn_gen_star_inputs_f32")
    return inputs, forcing


def box_masks_3d(params, size: int, device="cuda"):
    """Union of <= 3 axis-aligned boxes per sample as `source`, the six faces as `sink`: params (B, 19) int32 =
    (n, lo[3][3], hi[3][3]) with [box][axis (z, y, x)] half-open ranges.  Returns (source, sink, forcing)."""
    dev = _need_cuda(device)
    p = torch.as_tensor(params, dtype=torch.int32).to(dev).contiguous()
    B = p.shape[0]
    shp = (B, 1, size, size, size)
    src, sink, f = (torch.empty(shp, device=dev, dtype=torch.float32) for _ in range(3))
    with torch.cuda.device(dev):
        rc = L.lib().dn_gen_box_masks_3d_f32(C.c_void_p(p.data_ptr()), B, size, C.c_void_p(src.data_ptr()),
                                             C.c_void_p(sink.data_ptr()), C.c_void_p(f.data_ptr()), _stream(dev))
    L.check(rc, "dn_gen_box_masks_3d_f32")
    return src, sink, f


class _DeviceDataset(torch.utils.data.Dataset):
    n_samples = 0

    def __len__(self):
        return self.n_samples

    def __getitem__(self, index):
        inputs, forcing = self.batch([index])
        return inputs[0], forcing[0]


class KLSumStochastic(_DeviceDataset):
    """``KLSumStochastic(filename, domain_size=64, kl_terms=6)`` (klsum.py:10-46): ``filename`` = the .npy of Sobol
    coefficients (or the array itself).  Only the (n, 6) fp64 coefficients live on the device; fields are produced
    per batch."""

    def __init__(self, filename, domain_size=64, kl_terms=6, device="cuda", eta=0.5, omega=None):
        coeffs = np.load(filename) if isinstance(filename, (str, os.PathLike)) else np.asarray(filename)
        self.device = _need_cuda(device)
        self.coeffs = torch.as_tensor(coeffs[:, :kl_terms], dtype=torch.float64).to(self.device).contiguous()
        self.domain_size, self.kl_terms, self.eta, self.omega = domain_size, kl_terms, eta, omega
        self.n_samples = self.coeffs.shape[0]

    def batch(self, indices):
        idx = torch.as_tensor(indices, device=self.device, dtype=torch.long)
        return kl_inputs(self.coeffs[idx], self.domain_size, 2, self.eta, self.omega, self.device)


class ImageIMBack(_DeviceDataset):
    """``ImageIMBack(dirname, domain_size=64)`` (images.py:9-49).  ``dirname``: a directory of .png/.jpg/.bmp/.tiff
    files (decoded once on the host with PIL, kept on the device as bytes) or a (n, H, W) uint8 array."""

    def __init__(self, dirname, domain_size=64, device="cuda"):
        self.device = _need_cuda(device)
        if isinstance(dirname, (str, os.PathLike)):
            import PIL.Image
            imgs = []
            for fname in sorted(os.listdir(dirname)):
                ext = os.path.splitext(fname)[1]
                if ext not in (".png", ".jpg", ".bmp", ".tiff"):
                    raise ValueError("invalid extension; extension not supported")
                imgs.append(np.asarray(PIL.Image.open(os.path.join(dirname, fname)).convert("L")))
            images = np.stack(imgs)
        else:
            images = np.asarray(dirname)
        self.images = torch.as_tensor(images.astype(np.uint8)).to(self.device).contiguous()
        self.n_samples = self.images.shape[0]

    def batch(self, indices):
        idx = torch.as_tensor(indices, device=self.device, dtype=torch.long)
        return image_inputs(self.images[idx], self.device)


def read_voxel_config(config_name):
    """(bBoxMax, bBoxMin, numDiv, gridSize) of a VoxelConfig.txt (voxels.py:9-23)."""
    with open(config_name) as fh:
        fh.readline()
        bmin = np.array([float(v) for v in fh.readline().split()])
        bmax = np.array([float(v) for v in fh.readline().split()])
        num_div = np.array([int(v) for v in fh.readline().split()])
        grid = np.array([float(v) for v in fh.readline().split()])
    return bmax, bmin, num_div, grid


class VoxelIMBackRAW(_DeviceDataset):
    """``VoxelIMBackRAW(filename, domain_size=64)`` (voxels.py:31-61): ``filename`` + 'inouts.raw' / 'VoxelConfig.txt'."""

    def __init__(self, filename, domain_size=64, device="cuda", offset=32):
        self.device = _need_cuda(device)
        raw = np.fromfile(filename + "inouts.raw", dtype=np.uint8)
        _, _, num_div, _ = read_voxel_config(filename + "VoxelConfig.txt")
        self.inputs, self.forcing = voxel_inputs(torch.as_tensor(raw), num_div, domain_size, offset, self.device)
        self.n_samples = 100

    def batch(self, indices):
        n = len(indices)
        return self.inputs.expand(n, -1, -1, -1, -1), self.forcing.expand(n, -1, -1, -1, -1)
