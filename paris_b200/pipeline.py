"""Host-side mirror of the reference's backend-agnostic pipeline wrappers, over the C ABI.

The reference's hot loop (/root/reference/src/main.cpp:98-105) is

    p   = source.load_next()
    d_p = load(p)                      # src/loader.cpp:28-33
    weight(d_p, det_geo)               # src/weighting.cpp:32-45
    filter(d_p, det_geo)               # src/filtering.cpp:32-45
    backproject(d_p, v, offset, ...)   # src/backprojection.cpp:37-69
    ...
    sink.save(v)                       # copy_d2h

and the functions below have the same names, argument meaning and derived constants (computed in
float32 like the reference), so the parity tests read like that loop.  The product's C++ form of
this mirror is paris_b200/cpp (namespace paris::b200); this Python form exists for tests/ and
bench.py.  All arithmetic on the data happens in libparis_b200.so -- there is no CPU path here.
"""
from __future__ import annotations

import ctypes as C
import ctypes.util
from dataclasses import dataclass

import numpy as np

from . import capi
from .capi import Context, DetectorGeometry, Roi, VolumeGeometry

_libm = C.CDLL(ctypes.util.find_library("m") or "libm.so.6")
_libm.sinf.restype = C.c_float
_libm.sinf.argtypes = [C.c_float]
_libm.cosf.restype = C.c_float
_libm.cosf.argtypes = [C.c_float]

f32 = np.float32


@dataclass
class DeviceProjection:
    """projection<device buffer> of src/projection.h:31-46"""
    d_ptr: int
    dim_x: int
    dim_y: int
    idx: int = 0
    phi: float = 0.0


@dataclass
class DeviceVolume:
    """volume<device buffer> of src/volume.h:31-45"""
    d_ptr: int
    dim_x: int
    dim_y: int
    dim_z: int
    off: int = 0

    @property
    def dims(self):
        return (self.dim_x, self.dim_y, self.dim_z)

    @property
    def n_voxels(self):
        return self.dim_x * self.dim_y * self.dim_z


def weight_constants(det: DetectorGeometry):
    """h_min, v_min, d_sd as src/weighting.cpp:37-42 derives them (float32)."""
    n_row_f, n_col_f = f32(det.n_row), f32(det.n_col)
    l_row, l_col = f32(det.l_px_row), f32(det.l_px_col)
    h_min = f32(f32(det.delta_s) * l_row) - f32(f32(n_row_f * l_row) / f32(2))
    v_min = f32(f32(det.delta_t) * l_col) - f32(f32(n_col_f * l_col) / f32(2))
    d_sd = f32(abs(f32(det.d_so))) + f32(abs(f32(det.d_od)))
    return float(h_min), float(v_min), float(d_sd)


def angle_sin_cos(idx: int, det: DetectorGeometry, phi_deg: float | None = None):
    """sin/cos of the projection angle as src/backprojection.cpp:53-63 computes them (float32, libm)."""
    phi = f32(phi_deg) if phi_deg is not None else f32(f32(idx) * f32(det.delta_phi))
    phi = f32(phi * f32(f32(np.pi) / f32(180.0)))
    return float(_libm.sinf(float(phi))), float(_libm.cosf(float(phi)))


class Pipeline:
    """Per-context state the reference keeps in function-local statics: the filter table
    (thread_local static k, src/filtering.cpp:42)."""

    def __init__(self, ctx: Context, det: DetectorGeometry):
        self.ctx = ctx
        self.det = det
        self.filter_size = capi.filter_size(det.n_row)
        self._filter = None

    def close(self):
        if self._filter is not None:
            self.ctx.filter_destroy(self._filter)
            self._filter = None

    @property
    def filter_handle(self) -> int:
        if self._filter is None:
            self._filter = self.ctx.filter_create(self.filter_size, float(self.det.l_px_row))
        return self._filter

    # -- src/loader.cpp:28-33
    def load(self, h_proj: np.ndarray, idx: int = 0, phi: float = 0.0) -> DeviceProjection:
        dim_y, dim_x = h_proj.shape
        d = self.ctx.dev_alloc(dim_x * dim_y * 4)
        self.ctx.proj_h2d(h_proj, d, dim_x, dim_y)
        return DeviceProjection(d, dim_x, dim_y, idx, phi)

    def release(self, p: DeviceProjection):
        self.ctx.dev_free(p.d_ptr)

    def download(self, p: DeviceProjection) -> np.ndarray:
        out = np.empty((p.dim_y, p.dim_x), dtype=np.float32)
        self.ctx.proj_d2h(p.d_ptr, out, p.dim_x, p.dim_y)
        return out

    # -- src/make_volume.cpp:30-37
    def make_volume(self, dim_x: int, dim_y: int, dim_z: int) -> DeviceVolume:
        return DeviceVolume(self.ctx.volume_alloc(dim_x, dim_y, dim_z), dim_x, dim_y, dim_z)

    def save(self, v: DeviceVolume) -> np.ndarray:
        """sink.save: copy_d2h (src/sink.cpp:76-77); returns (dim_z, dim_y, dim_x)."""
        out = np.empty((v.dim_z, v.dim_y, v.dim_x), dtype=np.float32)
        self.ctx.vol_d2h(v.d_ptr, out, v.n_voxels)
        return out

    def free_volume(self, v: DeviceVolume):
        self.ctx.volume_free(v.d_ptr)

    # -- src/weighting.cpp:32-45
    def weight(self, p: DeviceProjection):
        h_min, v_min, d_sd = weight_constants(self.det)
        self.ctx.weight(p.d_ptr, p.dim_x, p.dim_y, h_min, v_min, d_sd, self.det.l_px_row, self.det.l_px_col)

    # -- src/filtering.cpp:32-45
    def filter(self, p: DeviceProjection):
        self.ctx.apply_filter(p.d_ptr, p.dim_x, p.dim_y, self.filter_handle, self.filter_size, self.det.n_col)

    def weight_filter(self, p: DeviceProjection):
        h_min, v_min, d_sd = weight_constants(self.det)
        self.ctx.weight_filter(p.d_ptr, p.dim_x, p.dim_y, h_min, v_min, d_sd, self.det.l_px_row, self.det.l_px_col,
                               self.filter_handle, self.filter_size)

    # -- src/backprojection.cpp:37-69
    def backproject(self, p: DeviceProjection, v: DeviceVolume, v_offset: int, vol_geo: VolumeGeometry,
                    enable_angles: bool = False, roi: Roi | None = None, fused_raw: bool = False):
        delta_s = float(f32(self.det.delta_s) * f32(self.det.l_px_row))
        delta_t = float(f32(self.det.delta_t) * f32(self.det.l_px_col))
        sn, cs = angle_sin_cos(p.idx, self.det, p.phi if enable_angles else None)
        self.ctx.backproject(p.d_ptr, p.dim_x, p.dim_y, v.d_ptr, v.dims, v_offset, self.det, vol_geo, roi, sn, cs,
                             delta_s, delta_t,
                             capi.BP_FUSE_WEIGHT_FILTER if fused_raw else 0,
                             self.filter_handle if fused_raw else None)

    # -- the loop of src/main.cpp:93-107 for one task (one slab)
    def reconstruct(self, stack: np.ndarray, vol_dims, vol_geo: VolumeGeometry, roi: Roi | None = None,
                    v_offset: int = 0, fused: bool = True, first_idx: int = 0, idx_stride: int = 1) -> np.ndarray:
        """stack: (n_proj, n_col, n_row) raw projections on the host.  fused=False runs the three
        contract stages one by one (weight, apply_filter, backproject); fused=True hands the raw
        projection to backproject with PARIS_B200_BP_FUSE_WEIGHT_FILTER."""
        v = self.make_volume(*vol_dims)
        for i in range(stack.shape[0]):
            d_p = self.load(stack[i], idx=first_idx + i * idx_stride)
            if not fused:
                self.weight(d_p)
                self.filter(d_p)
            self.backproject(d_p, v, v_offset, vol_geo, roi=roi, fused_raw=fused)
            self.release(d_p)
        out = self.save(v)
        self.free_volume(v)
        return out
