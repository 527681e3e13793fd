"""torch.autograd.Function wrappers over the C ABI (the drop-in seam for DiffNet's loss()).

  fem_energy(...)    fused Poisson energy loss, differentiable w.r.t. u (and nu in 2-D)
  fem_residual(...)  assembled-residual loss  sum(R^2), differentiable w.r.t. u
  gp_eval(...)       gauss_pt_evaluation{,_der_x,_der_y,_der_z}, differentiable w.r.t. its input

Everything runs on the CUDA stream torch considers current; there is no host sync inside and no
CPU path (CPU tensors raise).  See include/diffnet_fem.h for the C side.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L


@dataclass(frozen=True)
class Geometry:
    """Mesh description shared by every call (mirrors the reference's kwargs -> attributes,
    DiffNet/base.py:16-32, DiffNetFEM.py:42-51)."""
    nsd: int
    nx: int
    ny: int
    nz: int
    hx: float
    hy: float
    hz: float
    ngp_1d: int = 2

    @property
    def spatial(self) -> Tuple[int, ...]:
        return (self.ny, self.nx) if self.nsd == 2 else (self.nz, self.ny, self.nx)

    @property
    def elems(self) -> Tuple[int, ...]:
        return tuple(s - 1 for s in self.spatial)


# ------------------------------------------------------------------------------ marshalling
def _require_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise L.DiffNetFEMError(
            f"{name} is on {t.device}: the FEM ops are CUDA (sm_100a) only, there is no CPU fallback")


def _canon(t: torch.Tensor, geom: Geometry, name: str) -> torch.Tensor:
    """View `t` as (Bt, *spatial) with x contiguous, fp32.  Accepted layouts: bare spatial,
    (B, *spatial) and (B, 1, *spatial) -- what the reference's where()/conv broadcasting accepts
    (solve_in_object_3d.py:198-199 passes a bare (D,H,W) parameter)."""
    _require_cuda(t, name)
    sp = geom.spatial
    nsd = geom.nsd
    if t.dtype != torch.float32:
        if t.dtype in (torch.bool, torch.uint8):
            t = t.to(torch.float32)
        else:
            raise L.DiffNetFEMError(f"{name} must be float32 (got {t.dtype})")
    if t.dim() == nsd + 2:
        if t.shape[1] != 1:
            raise L.DiffNetFEMError(f"{name}: expected one channel, got shape {tuple(t.shape)}")
        t = t[:, 0]
    elif t.dim() == nsd:
        t = t.unsqueeze(0)
    elif t.dim() != nsd + 1:
        raise L.DiffNetFEMError(f"{name}: bad shape {tuple(t.shape)} for a {nsd}-D nodal field")
    if tuple(t.shape[1:]) != sp:
        raise L.DiffNetFEMError(f"{name}: spatial shape {tuple(t.shape[1:])} != mesh nodes {sp}")
    if t.stride(-1) != 1:
        t = t.contiguous()
    return t


def _field(t: Optional[torch.Tensor], B: int, nsd: int) -> L.dn_field:
    if t is None:
        return L.dn_field(None, 0, 0, 0)
    bt = t.shape[0]
    if bt != B and bt != 1:
        raise L.DiffNetFEMError(f"batch {bt} does not broadcast to {B}")
    sb = t.stride(0) if (bt == B and B > 1) else 0
    if nsd == 2:
        return L.dn_field(t.data_ptr(), sb, 0, t.stride(1))
    return L.dn_field(t.data_ptr(), sb, t.stride(1), t.stride(2))


def _geom_struct(geom: Geometry, B: int, z_own=None, mean_count=0.0) -> L.dn_geom:
    zo = z_own or (0, 0)
    return L.dn_geom(geom.nsd, B, geom.nx, geom.ny, geom.nz if geom.nsd == 3 else 1, geom.ngp_1d,
                     geom.hx, geom.hy, geom.hz if geom.nsd == 3 else 0.0, int(zo[0]), int(zo[1]),
                     float(mean_count))


def _new_out(shape, device, dtype=torch.float32) -> torch.Tensor:
    """Output buffer.  With DN_POISON_OUTPUTS=1 (set by the test-suite) it is NaN-filled first so
    that a node the kernel failed to write can never pass a parity check by accident."""
    if os.environ.get("DN_POISON_OUTPUTS"):
        return torch.full(shape, float("nan"), dtype=dtype, device=device)
    return torch.empty(shape, dtype=dtype, device=device)


_workspaces = {}
_ws_bytes = {}


class _on_device:
    """`with torch.cuda.device(dev)` costs ~10 us per call; skip it when `dev` is already current
    (the one-process-per-GPU case)."""

    def __init__(self, dev):
        self.ctx = None if torch.cuda.current_device() == dev.index else torch.cuda.device(dev)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *a):
        if self.ctx is not None:
            self.ctx.__exit__(*a)


def _workspace_for(lib, g, geom, B, device, stream_ptr):
    key = (geom.nsd, B, geom.nx, geom.ny, geom.nz)
    n = _ws_bytes.get(key)
    if n is None:
        n = _ws_bytes[key] = lib.dn_fem_workspace_bytes(C.byref(g))
    return _workspace(device, stream_ptr, n)


def _workspace(device: torch.device, stream_ptr: int, nbytes: int) -> torch.Tensor:
    """Zero-initialised scratch (ticket counter + per-CTA partials), one per device and stream;
    the kernels leave it ready for the next call."""
    key = (device.index, stream_ptr)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(max(nbytes, 1 << 16), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def _masks_struct(dirichlet, geom: Geometry, B: int, keep):
    n = len(dirichlet)
    if n > L.DN_MAX_MASKS:
        raise L.DiffNetFEMError(f"at most {L.DN_MAX_MASKS} Dirichlet conditions per call (got {n})")
    arr = (L.dn_mask * max(n, 1))()
    for i, (mask, value) in enumerate(dirichlet):
        m = _canon(mask, geom, f"dirichlet[{i}].mask")
        keep.append(m)
        arr[i].mask = _field(m, B, geom.nsd)
        if torch.is_tensor(value):
            v = _canon(value, geom, f"dirichlet[{i}].value")
            keep.append(v)
            arr[i].value_field = _field(v, B, geom.nsd)
            arr[i].value = 0.0
        else:
            arr[i].value_field = L.dn_field(None, 0, 0, 0)
            arr[i].value = float(value)
    return arr, n


def _batch_of(*ts) -> int:
    B = 1
    for t in ts:
        if t is not None:
            B = max(B, t.shape[0])
    return B


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _check_f_gp(geom: Geometry, f_gp) -> torch.Tensor:
    _require_cuda(f_gp, "f_gp")
    ngp = geom.ngp_1d ** geom.nsd
    fg = f_gp if f_gp.dim() == geom.nsd + 2 else f_gp.unsqueeze(0)
    if tuple(fg.shape[1:]) != (ngp,) + geom.elems or fg.dtype != torch.float32:
        raise L.DiffNetFEMError(
            f"f_gp must be float32 (B|1, {ngp}, {geom.elems}); got {tuple(f_gp.shape)} {f_gp.dtype}")
    return fg


def load_vector(geom: Geometry, f_gp: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Assembled load vector b (Bf, *nodes) of a forcing given at the Gauss points (Bf = f_gp's batch, 1 when it is
    shared by the batch): b_a = sum over elements and Gauss points of w_g N_a(g) f_gp.  The forcing term of the
    reference's f_gp form (e8_2d_poisson_mms.py:154-175) is linear in u, so sum_g w_g f_g u_g = sum_a b_a u_a:
    pass ``b`` as ``f`` with ``load_vector=True`` and the fused kernels stream 4 bytes per node instead of
    4 * ngp bytes per element."""
    fg = _check_f_gp(geom, f_gp).contiguous()
    Bf = fg.shape[0]
    dev = fg.device
    if out is None:
        out = _new_out((Bf,) + geom.spatial, dev)
    g = _geom_struct(geom, Bf)
    ffg = L.dn_field(fg.data_ptr(), fg.stride(0) if Bf > 1 else 0, 0, 0)
    with _on_device(dev):
        rc = L.lib().dn_fem_load_vector_f32(C.byref(ffg), C.byref(g), C.c_void_p(out.data_ptr()),
                                            C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    L.check(rc, "dn_fem_load_vector_f32")
    return out


# f_gp -> load vector, remembered per tensor OBJECT and version counter (an in-place update of f_gp re-assembles; a
# new tensor at a recycled address is a different object and misses).  Small: a training loop has one or two f_gp.
_LV_CACHE = []          # [(weakref(f_gp), version, geom key, b)]
_LV_CACHE_MAX = 8
USE_LOAD_VECTOR = os.environ.get("DN_LOAD_VECTOR", "1") != "0"


def _cached_load_vector(geom: Geometry, f_gp: torch.Tensor) -> torch.Tensor:
    import weakref
    key = (geom.nsd, geom.spatial, geom.ngp_1d)
    for i, (ref, ver, k, b) in enumerate(_LV_CACHE):
        if ref() is f_gp and k == key:
            if ver == f_gp._version:
                return b
            _LV_CACHE.pop(i)
            break
    _LV_CACHE[:] = [e for e in _LV_CACHE if e[0]() is not None][-(_LV_CACHE_MAX - 1):]
    b = load_vector(geom, f_gp)
    _LV_CACHE.append((weakref.ref(f_gp), f_gp._version, key, b))
    return b


def energy_raw(geom: Geometry, u, nu=None, f=None, f_gp=None, dirichlet=(), nu_zero_mask=None,
               c_k=1.0, c_f=1.0, scale=1.0, reduction="mean", want_grad=True, want_grad_nu=False,
               z_own=None, mean_count=0.0, want_double=False, load_vector=False):
    """One fused launch.  Returns (loss 0-dim fp32, grad_u (B,*spatial) or None,
    grad_nu or None[, loss fp64 0-dim]).

    ``load_vector=True``: ``f`` is an assembled load vector (``load_vector()``), not a nodal source.  A call with
    ``f_gp`` takes that route by itself (assembling b once per f_gp tensor and version) whenever the streaming
    kernels can run the launch, and reads f_gp in the general kernels otherwise."""
    if (f_gp is not None and f is None and USE_LOAD_VECTOR and not load_vector and z_own is None
            and not want_grad_nu):
        _check_f_gp(geom, f_gp)
        try:
            return energy_raw(geom, u, nu=nu, f=_cached_load_vector(geom, f_gp), dirichlet=dirichlet,
                              nu_zero_mask=nu_zero_mask, c_k=c_k, c_f=c_f, scale=scale, reduction=reduction,
                              want_grad=want_grad, want_grad_nu=False, mean_count=mean_count,
                              want_double=want_double, load_vector=True)
        except L.DiffNetFEMError as e:
            if e.code != L.DN_ENOSTREAM:
                raise                      # anything else is a real error; DN_ENOSTREAM: the general kernels read f_gp
    keep = []
    uc = _canon(u, geom, "u")
    nuc = _canon(nu, geom, "nu") if nu is not None else None
    fc = _canon(f, geom, "f") if f is not None else None
    nzm = _canon(nu_zero_mask, geom, "nu_zero_mask") if nu_zero_mask is not None else None
    fg = None
    if f_gp is not None:
        fg = _check_f_gp(geom, f_gp).contiguous()
    if load_vector and (fc is None or fg is not None):
        raise L.DiffNetFEMError("load_vector=True: pass the assembled vector as f (and no f_gp)")
    masks_t = [m for m, _ in dirichlet] + [v for _, v in dirichlet if torch.is_tensor(v)]
    B = _batch_of(uc, nuc, fc, nzm, fg, *[_canon(m, geom, "mask") for m in masks_t])
    marr, nm = _masks_struct(dirichlet, geom, B, keep)
    dev = uc.device
    g = _geom_struct(geom, B, z_own, mean_count)
    cs = L.dn_consts(float(c_k), float(c_f), float(scale), 0 if reduction == "mean" else 1,
                     L.DN_F_LOAD_VECTOR if load_vector else 0)
    if reduction not in ("mean", "sum"):
        raise L.DiffNetFEMError("reduction must be 'mean' or 'sum'")
    grad = _new_out((B,) + geom.spatial, dev) if want_grad else None
    grad_nu = _new_out((B,) + geom.spatial, dev) if want_grad_nu else None
    loss = _new_out((), dev)
    loss64 = torch.empty((), dtype=torch.float64, device=dev) if want_double else None
    lib = L.lib()
    stream = torch.cuda.current_stream(dev).cuda_stream
    with _on_device(dev):
        ws = _workspace_for(lib, g, geom, B, dev, stream)
        fu, fnu, ff = _field(uc, B, geom.nsd), _field(nuc, B, geom.nsd), _field(fc, B, geom.nsd)
        ffg = L.dn_field(None, 0, 0, 0)
        if fg is not None:
            ffg = L.dn_field(fg.data_ptr(), fg.stride(0) if (fg.shape[0] == B and B > 1) else 0, 0, 0)
        fnz = _field(nzm, B, geom.nsd)
        fn = lib.dn_fem_energy_2d_f32 if geom.nsd == 2 else lib.dn_fem_energy_3d_f32
        rc = fn(C.byref(fu), C.byref(fnu) if nuc is not None else None,
                C.byref(ff) if fc is not None else None, C.byref(ffg) if fg is not None else None,
                marr, nm, C.byref(fnz) if nzm is not None else None, C.byref(g), C.byref(cs),
                _ptr(grad), _ptr(grad_nu), C.c_void_p(ws.data_ptr()), ws.numel(), _ptr(loss64),
                _ptr(loss), C.c_void_p(stream))
    L.check(rc, f"dn_fem_energy_{geom.nsd}d_f32")
    out = (loss, grad, grad_nu)
    return out + (loss64,) if want_double else out


class PreparedEnergy:
    """A fused energy call bound to fixed tensors: all marshalling (views, strides, ctypes
    structs, workspace, output buffers) is done once; ``__call__`` is one C call (~5 us of host
    time instead of ~50).  For loops that evaluate the loss on the SAME storage every iteration:
    u-as-parameter solves (solve_in_object_3d.py, LBFGS closures), the z-slab steps, benchmarks.

    The returned ``(loss, grad)`` are the SAME tensors on every call (overwritten in place):
    consume or copy them before calling again.  In-place updates of the bound tensors are seen
    (pointers, not values, are captured); re-prepare if a tensor is reallocated."""

    def __init__(self, geom: Geometry, u, nu=None, f=None, f_gp=None, dirichlet=(), nu_zero_mask=None,
                 c_k=1.0, c_f=1.0, scale=1.0, reduction="mean", z_own=None, mean_count=0.0, link=None):
        """``link``: a ``_lib.dn_slab_link`` (3-D only, see include/diffnet_fem.h) -- the launch then also
        exchanges the z-halo planes of u with the neighbouring ranks and pushes the loss partial."""
        if reduction not in ("mean", "sum"):
            raise L.DiffNetFEMError("reduction must be 'mean' or 'sum'")
        if link is not None and (geom.nsd != 3 or f_gp is not None):
            raise L.DiffNetFEMError("a linked (z-slab) call is 3-D with a nodal source term")
        self._link = link
        self.geom = geom
        keep = []
        uc = _canon(u, geom, "u")
        nuc = _canon(nu, geom, "nu") if nu is not None else None
        fc = _canon(f, geom, "f") if f is not None else None
        nzm = _canon(nu_zero_mask, geom, "nu_zero_mask") if nu_zero_mask is not None else None
        # the call is bound to STORAGE: a tensor that had to be copied (dtype conversion, x not
        # contiguous) would be read from a stale private copy on every later call
        bound = [("u", u, uc), ("nu", nu, nuc), ("f", f, fc), ("nu_zero_mask", nu_zero_mask, nzm)]
        bound += [(f"dirichlet[{i}].mask", m, _canon(m, geom, "mask")) for i, (m, _) in enumerate(dirichlet)]
        bound += [(f"dirichlet[{i}].value", v, _canon(v, geom, "value")) for i, (_, v) in enumerate(dirichlet)
                  if torch.is_tensor(v)]
        for nm, orig, canon in bound:
            if orig is not None and canon.data_ptr() != orig.data_ptr():
                raise L.DiffNetFEMError(
                    f"prepare_energy: {nm} would have to be copied (dtype {orig.dtype}, strides {tuple(orig.stride())}); "
                    "a prepared call binds storage -- pass float32 tensors whose x axis is contiguous")
        fg = None
        if f_gp is not None:
            fg = _check_f_gp(geom, f_gp)
            if not fg.is_contiguous():
                raise L.DiffNetFEMError("prepare_energy: f_gp must be contiguous (a prepared call binds storage)")
            if f is not None:
                raise L.DiffNetFEMError("give f or f_gp, not both")
        masks_t = [m for m, _ in dirichlet] + [v for _, v in dirichlet if torch.is_tensor(v)]
        B = _batch_of(uc, nuc, fc, nzm, fg, *[_canon(m, geom, "mask") for m in masks_t])
        self._marr, self._nm = _masks_struct(dirichlet, geom, B, keep)
        self.device = dev = uc.device
        self._g = _geom_struct(geom, B, z_own, mean_count)
        self._cs = L.dn_consts(float(c_k), float(c_f), float(scale), 0 if reduction == "mean" else 1, 0)
        self.grad = _new_out((B,) + geom.spatial, dev)
        self.loss = _new_out((), dev)
        self._fu, self._fnu, self._ff = _field(uc, B, geom.nsd), _field(nuc, B, geom.nsd), _field(fc, B, geom.nsd)
        self._ffg = L.dn_field(None, 0, 0, 0)
        if fg is not None:
            self._ffg = L.dn_field(fg.data_ptr(), fg.stride(0) if (fg.shape[0] == B and B > 1) else 0, 0, 0)
        self._fnz = _field(nzm, B, geom.nsd)
        self._has = (nuc is not None, fc is not None, fg is not None, nzm is not None)
        # f_gp: the assembled load vector is what the streaming kernels read (re-assembled when f_gp's version
        # counter moves); the first call falls back to the general kernels for good if they refuse the launch
        self._lv = None
        if fg is not None and link is None and z_own is None and USE_LOAD_VECTOR:
            bvec = load_vector(geom, fg)
            self._lv = (fg, fg._version, bvec, _field(bvec, B, geom.nsd),
                        L.dn_consts(float(c_k), float(c_f), float(scale), 0 if reduction == "mean" else 1,
                                    L.DN_F_LOAD_VECTOR))
        self._keep = (keep, uc, nuc, fc, nzm, fg)            # the views the pointers refer to
        lib = L.lib()
        self._fn = lib.dn_fem_energy_2d_f32 if geom.nsd == 2 else lib.dn_fem_energy_3d_f32
        self._B = B
        self._ws = {}

    def __call__(self):
        dev = self.device
        stream = torch.cuda.current_stream(dev).cuda_stream
        ws = self._ws.get(stream)
        with _on_device(dev):
            if ws is None:
                ws = self._ws[stream] = _workspace_for(L.lib(), self._g, self.geom, self._B, dev, stream)
            hn, hf, hg, hz = self._has
            if self._link is not None:
                rc = L.lib().dn_fem_energy_3d_linked_f32(
                    C.byref(self._fu), C.byref(self._fnu) if hn else None, C.byref(self._ff) if hf else None,
                    self._marr, self._nm, C.byref(self._fnz) if hz else None, C.byref(self._g), C.byref(self._cs),
                    C.byref(self._link), C.c_void_p(self.grad.data_ptr()), C.c_void_p(ws.data_ptr()), ws.numel(),
                    None, C.c_void_p(self.loss.data_ptr()), C.c_void_p(stream))
                L.check(rc, "dn_fem_energy_3d_linked_f32")
                return self.loss, self.grad
            if self._lv is not None:
                fg, ver, bvec, fb, cs = self._lv
                if fg._version != ver:
                    load_vector(self.geom, fg, out=bvec)
                    self._lv = (fg, fg._version, bvec, fb, cs)
                rc = self._fn(C.byref(self._fu), C.byref(self._fnu) if hn else None, C.byref(fb), None, self._marr,
                              self._nm, C.byref(self._fnz) if hz else None, C.byref(self._g), C.byref(cs),
                              C.c_void_p(self.grad.data_ptr()), None, C.c_void_p(ws.data_ptr()), ws.numel(), None,
                              C.c_void_p(self.loss.data_ptr()), C.c_void_p(stream))
                if rc != L.DN_ENOSTREAM:
                    L.check(rc, f"dn_fem_energy_{self.geom.nsd}d_f32")
                    return self.loss, self.grad
                self._lv = None
            rc = self._fn(C.byref(self._fu), C.byref(self._fnu) if hn else None, C.byref(self._ff) if hf else None,
                          C.byref(self._ffg) if hg else None, self._marr, self._nm,
                          C.byref(self._fnz) if hz else None, C.byref(self._g), C.byref(self._cs),
                          C.c_void_p(self.grad.data_ptr()), None, C.c_void_p(ws.data_ptr()), ws.numel(), None,
                          C.c_void_p(self.loss.data_ptr()), C.c_void_p(stream))
        L.check(rc, f"dn_fem_energy_{self.geom.nsd}d_f32")
        return self.loss, self.grad


def residual_raw(geom: Geometry, u, nu=None, f=None, dirichlet=(), jac=1.0, apply_masks_to_input=True):
    """R = jac*(K(nu) u' - F(f)) masked, and sum(R^2).  Returns (loss 0-dim, R (B,*spatial))."""
    keep = []
    uc = _canon(u, geom, "u")
    nuc = _canon(nu, geom, "nu") if nu is not None else None
    fc = _canon(f, geom, "f") if f is not None else None
    B = _batch_of(uc, nuc, fc, *[_canon(m, geom, "mask") for m, _ in dirichlet])
    marr, nm = _masks_struct(dirichlet, geom, B, keep)
    dev = uc.device
    g = _geom_struct(geom, B)
    R = _new_out((B,) + geom.spatial, dev)
    loss = _new_out((), dev)
    lib = L.lib()
    stream = torch.cuda.current_stream(dev).cuda_stream
    with _on_device(dev):
        ws = _workspace_for(lib, g, geom, B, dev, stream)
        fu, fnu, ff = _field(uc, B, geom.nsd), _field(nuc, B, geom.nsd), _field(fc, B, geom.nsd)
        fn = lib.dn_fem_residual_2d_f32 if geom.nsd == 2 else lib.dn_fem_residual_3d_f32
        rc = fn(C.byref(fu), C.byref(fnu) if nuc is not None else None,
                C.byref(ff) if fc is not None else None, marr, nm, 1 if apply_masks_to_input else 0,
                C.byref(g), float(jac), _ptr(R), C.c_void_p(ws.data_ptr()), ws.numel(), None,
                _ptr(loss), C.c_void_p(stream))
    L.check(rc, f"dn_fem_residual_{geom.nsd}d_f32")
    return loss, R


def scale_inplace_(x: torch.Tensor, factor: torch.Tensor) -> torch.Tensor:
    """x *= factor (0-dim device tensor), free when factor == 1 (checked on the device)."""
    _require_cuda(x, "x")
    assert x.is_contiguous() and x.dtype == torch.float32
    fac = factor.detach().to(device=x.device, dtype=torch.float32).reshape(())
    stream = torch.cuda.current_stream(x.device).cuda_stream
    with _on_device(x.device):
        rc = L.lib().dn_scale_inplace_f32(C.c_void_p(x.data_ptr()), x.numel(),
                                          C.c_void_p(fac.data_ptr()), C.c_void_p(stream))
    L.check(rc, "dn_scale_inplace_f32")
    return x


def _like_input(grad_b: torch.Tensor, ref: torch.Tensor, geom: Geometry) -> torch.Tensor:
    """Reduce/reshape a dense (B,*spatial) gradient to the shape of the tensor the user passed."""
    if ref.dim() == geom.nsd:                       # bare field broadcast over the batch
        return grad_b.sum(0) if grad_b.shape[0] > 1 else grad_b[0]
    bt = ref.shape[0]
    g = grad_b
    if bt == 1 and g.shape[0] > 1:
        g = g.sum(0, keepdim=True)
    return g.reshape(ref.shape)


# ------------------------------------------------------------------------------ autograd
class FEMEnergyFunction(torch.autograd.Function):
    """loss = energy(u, nu, ...); backward returns grad_output * dloss/du (and dloss/dnu).

    Forward AND gradient come from the single fused launch in forward(); backward only scales
    (in place, skipped on the device when grad_output == 1).  First-order only
    (once_differentiable), which is all Adam / LBFGS need.

    The gradient buffer is handed to autograd, not copied, and scaling it in place consumes it: a
    SECOND backward through the same node (``retain_graph=True``, or ``autograd.grad`` followed by
    ``backward``) raises instead of silently returning ``grad_output**2 * dL/du`` -- re-run the
    forward (one launch) to differentiate again."""

    @staticmethod
    def forward(ctx, u, nu, geom, f, f_gp, dirichlet, nu_zero_mask, c_k, c_f, scale, reduction,
                z_own, mean_count):
        need_u = ctx.needs_input_grad[0]
        need_nu = nu is not None and ctx.needs_input_grad[1]
        loss, grad, grad_nu = energy_raw(
            geom, u.detach(), None if nu is None else nu.detach(), f, f_gp, dirichlet,
            nu_zero_mask, c_k, c_f, scale, reduction, want_grad=need_u or need_nu,
            want_grad_nu=need_nu, z_own=z_own, mean_count=mean_count)
        ctx.geom = geom
        ctx.u_ref = u if need_u else None
        ctx.nu_ref = nu if need_nu else None
        ctx.save_for_backward(*[t for t in (grad if need_u else None, grad_nu) if t is not None])
        ctx.has = (need_u, need_nu)
        ctx.consumed = False
        return loss

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gout):
        if ctx.consumed:
            raise RuntimeError(
                "FEMEnergyFunction: backward was already run through this node; its gradient buffer was scaled "
                "in place and handed to autograd.  Evaluate the loss again (one fused launch) for another backward.")
        ctx.consumed = True
        saved = list(ctx.saved_tensors)
        gu = gnu = None
        if ctx.has[0]:
            gu = _like_input(scale_inplace_(saved.pop(0), gout), ctx.u_ref, ctx.geom)
        if ctx.has[1]:
            gnu = _like_input(scale_inplace_(saved.pop(0), gout), ctx.nu_ref, ctx.geom)
        return (gu, gnu) + (None,) * 11


class FEMResidualFunction(torch.autograd.Function):
    """loss = sum(R^2); dL/du = mask( K(nu) (2R) ) -- one more pass of the same operator
    (K is symmetric; R is already zero on the Dirichlet nodes)."""

    @staticmethod
    def forward(ctx, u, geom, nu, f, dirichlet, jac):
        loss, R = residual_raw(geom, u.detach(), nu, f, dirichlet, jac, True)
        ctx.geom, ctx.nu, ctx.dirichlet, ctx.jac, ctx.u_ref = geom, nu, dirichlet, jac, u
        ctx.save_for_backward(R)
        return loss

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gout):
        (R,) = ctx.saved_tensors
        _, KR = residual_raw(ctx.geom, R, ctx.nu, None, ctx.dirichlet, ctx.jac, False)
        g = KR.mul_(2.0)
        g = scale_inplace_(g, gout)
        return (_like_input(g, ctx.u_ref, ctx.geom),) + (None,) * 5


_WHICH = {"N": 0, "dx": 1, "dy": 2, "dz": 3}


def _gp_raw(geom: Geometry, t: torch.Tensor, which: int) -> torch.Tensor:
    tc = _canon(t, geom, "tensor")
    B = tc.shape[0]
    g = _geom_struct(geom, B)
    out = _new_out((B, geom.ngp_1d ** geom.nsd) + geom.elems, tc.device)
    fld = _field(tc, B, geom.nsd)
    stream = torch.cuda.current_stream(tc.device).cuda_stream
    lib = L.lib()
    fn = lib.dn_fem_gp_eval_2d_f32 if geom.nsd == 2 else lib.dn_fem_gp_eval_3d_f32
    with _on_device(tc.device):
        rc = fn(C.byref(fld), C.byref(g), which, _ptr(out), C.c_void_p(stream))
    L.check(rc, "dn_fem_gp_eval")
    return out


def _gp_adj_raw(geom: Geometry, gout: torch.Tensor, which: int) -> torch.Tensor:
    """Transpose of `_gp_raw`: cotangent (B, ngp, elems) -> nodes (B, *spatial)."""
    gout = gout.contiguous()
    B = gout.shape[0]
    g = _geom_struct(geom, B)
    gin = _new_out((B,) + geom.spatial, gout.device)
    stream = torch.cuda.current_stream(gout.device).cuda_stream
    lib = L.lib()
    fn = lib.dn_fem_gp_eval_adj_2d_f32 if geom.nsd == 2 else lib.dn_fem_gp_eval_adj_3d_f32
    with _on_device(gout.device):
        rc = fn(_ptr(gout), C.byref(g), which, _ptr(gin), C.c_void_p(stream))
    L.check(rc, "dn_fem_gp_eval_adj")
    return gin


class GaussPointEvalFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t, geom, which):
        ctx.geom, ctx.which, ctx.t_ref = geom, which, t
        return _gp_raw(geom, t.detach(), which)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gout):
        return _like_input(_gp_adj_raw(ctx.geom, gout, ctx.which), ctx.t_ref, ctx.geom), None, None


def _gp_multi_raw(geom: Geometry, t: torch.Tensor, whichs: Sequence[int]):
    """Several tables in one call (dn_fem_gp_eval_multi_*): a tuple of (B, ngp, elems) tensors."""
    tc = _canon(t, geom, "tensor")
    B = tc.shape[0]
    g = _geom_struct(geom, B)
    outs = [_new_out((B, geom.ngp_1d ** geom.nsd) + geom.elems, tc.device) for _ in whichs]
    fld = _field(tc, B, geom.nsd)
    stream = torch.cuda.current_stream(tc.device).cuda_stream
    lib = L.lib()
    fn = lib.dn_fem_gp_eval_multi_2d_f32 if geom.nsd == 2 else lib.dn_fem_gp_eval_multi_3d_f32
    n = len(whichs)
    wh = (C.c_int * n)(*whichs)
    ptrs = (C.c_void_p * n)(*[o.data_ptr() for o in outs])
    with _on_device(tc.device):
        rc = fn(C.byref(fld), C.byref(g), n, wh, ptrs, C.c_void_p(stream))
    L.check(rc, "dn_fem_gp_eval_multi")
    return tuple(outs)


class GaussPointEvalMultiFunction(torch.autograd.Function):
    """gauss_pt_evaluation + its derivatives of ONE nodal field in one launch; backward sums the
    cotangents of all tables in one adjoint launch."""

    @staticmethod
    def forward(ctx, t, geom, whichs):
        ctx.geom, ctx.whichs, ctx.t_ref = geom, tuple(whichs), t
        return _gp_multi_raw(geom, t.detach(), whichs)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, *gouts):
        geom = ctx.geom
        used = [(w, g.contiguous()) for w, g in zip(ctx.whichs, gouts) if g is not None]
        if not used:
            return None, None, None
        B = used[0][1].shape[0]
        g = _geom_struct(geom, B)
        dev = used[0][1].device
        gin = _new_out((B,) + geom.spatial, dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        lib = L.lib()
        fn = lib.dn_fem_gp_eval_multi_adj_2d_f32 if geom.nsd == 2 else lib.dn_fem_gp_eval_multi_adj_3d_f32
        n = len(used)
        wh = (C.c_int * n)(*[w for w, _ in used])
        ptrs = (C.c_void_p * n)(*[t.data_ptr() for _, t in used])
        with _on_device(dev):
            rc = fn(ptrs, C.byref(g), n, wh, _ptr(gin), C.c_void_p(stream))
        L.check(rc, "dn_fem_gp_eval_multi_adj")
        return _like_input(gin, ctx.t_ref, geom), None, None


class GaussPointEvalGeneralFunction(torch.autograd.Function):
    """gauss_pt_eval for any tensor-product Lagrange basis / rule in 1-3 dimensions
    (dn_fem_gp_eval_general_f32): the degree 2 / 3 bases and the 1-D surface stencils."""

    @staticmethod
    def forward(ctx, t, nsd, nb, ng, factors):
        _require_cuda(t, "tensor")
        if t.dtype != torch.float32:
            raise L.DiffNetFEMError(f"tensor must be float32 (got {t.dtype})")
        tc = t.detach()
        if tc.dim() == nsd + 2:
            if tc.shape[1] != 1:
                raise L.DiffNetFEMError(f"expected one channel, got shape {tuple(tc.shape)}")
            tc = tc[:, 0]
        elif tc.dim() == nsd:
            tc = tc.unsqueeze(0)
        elif tc.dim() != nsd + 1:
            raise L.DiffNetFEMError(f"bad shape {tuple(tc.shape)} for a {nsd}-D nodal field")
        if tc.stride(-1) != 1:
            tc = tc.contiguous()
        B = tc.shape[0]
        sp = tuple(tc.shape[1:])                                   # ([nz, ny,] nx)
        if any((n - 1) % (nb - 1) for n in sp):
            raise L.DiffNetFEMError(f"nodes {sp}: (n - 1) must be a multiple of the basis degree {nb - 1}")
        nel = tuple((n - 1) // (nb - 1) for n in sp)
        n3 = (1,) * (3 - nsd) + sp                                 # (nz, ny, nx)
        fac = np.ascontiguousarray(np.asarray(factors, dtype=np.float32).reshape(nsd, ng, nb))
        out = _new_out((B, ng ** nsd) + nel, tc.device)
        fld = L.dn_field(tc.data_ptr(), tc.stride(0), tc.stride(1) if nsd == 3 else 0,
                         tc.stride(nsd - 1) if nsd >= 2 else 0)
        stream = torch.cuda.current_stream(tc.device).cuda_stream
        with _on_device(tc.device):
            rc = L.lib().dn_fem_gp_eval_general_f32(C.byref(fld), nsd, B, n3[2], n3[1], n3[0], nb, ng,
                                                    fac.ctypes.data_as(C.POINTER(C.c_float)), _ptr(out), C.c_void_p(stream))
        L.check(rc, "dn_fem_gp_eval_general_f32")
        ctx.meta = (nsd, nb, ng, fac, n3, B, sp)
        ctx.t_shape = tuple(t.shape)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gout):
        nsd, nb, ng, fac, n3, B, sp = ctx.meta
        gout = gout.contiguous()
        gin = _new_out((B,) + sp, gout.device)
        stream = torch.cuda.current_stream(gout.device).cuda_stream
        with _on_device(gout.device):
            rc = L.lib().dn_fem_gp_eval_general_adj_f32(_ptr(gout), nsd, B, n3[2], n3[1], n3[0], nb, ng,
                                                        fac.ctypes.data_as(C.POINTER(C.c_float)), _ptr(gin), C.c_void_p(stream))
        L.check(rc, "dn_fem_gp_eval_general_adj_f32")
        return gin.reshape(ctx.t_shape), None, None, None, None


def gp_eval_general(t: torch.Tensor, nsd: int, nbf_1d: int, ngp_1d: int, factors) -> torch.Tensor:
    """``factors[d][g][b]``, d = 0 (x) .. nsd-1: the 1-D basis value (or derivative * 2/h) of local node b at
    Gauss point g.  Returns (B, ngp_1d^nsd, elements...)."""
    return GaussPointEvalGeneralFunction.apply(t, nsd, nbf_1d, ngp_1d, factors)


# ------------------------------------------------------------------------------ public functional API
def fem_energy(geom: Geometry, u, nu=None, f=None, f_gp=None, dirichlet: Sequence = (),
               nu_zero_mask=None, c_k=1.0, c_f=1.0, scale=1.0, reduction="mean", z_own=None,
               mean_count=0.0) -> torch.Tensor:
    """Fused Poisson energy loss, differentiable w.r.t. u and (2-D) nu."""
    return FEMEnergyFunction.apply(u, nu, geom, f, f_gp, tuple(dirichlet), nu_zero_mask, c_k, c_f,
                                   scale, reduction, z_own, mean_count)


def fem_energy_and_grad(geom: Geometry, u, **kw):
    """(loss, dloss/du) from ONE launch, outside autograd (LBFGS closures, benchmarks)."""
    loss, grad, _ = energy_raw(geom, u, **kw)
    return loss, grad


def fem_residual(geom: Geometry, u, nu=None, f=None, dirichlet: Sequence = (), jac=1.0) -> torch.Tensor:
    return FEMResidualFunction.apply(u, geom, nu, f, tuple(dirichlet), jac)


def gp_eval(geom: Geometry, t: torch.Tensor, which: str = "N") -> torch.Tensor:
    return GaussPointEvalFunction.apply(t, geom, _WHICH[which])


def gp_eval_multi(geom: Geometry, t: torch.Tensor, which: Sequence[str] = ("N", "dx", "dy")):
    """(gauss_pt_evaluation(t), ..._der_x(t), ...) for the listed tables from ONE call (and one backward pass over all cotangents)."""
    whichs = tuple(_WHICH[w] for w in which)
    if not 1 <= len(whichs) <= 4 or (geom.nsd == 2 and 3 in whichs):
        raise ValueError(f"which={which!r}: 1..4 tables out of N, dx, dy" + (", dz" if geom.nsd == 3 else ""))
    return GaussPointEvalMultiFunction.apply(t, geom, whichs)
