"""ctypes binding of libparis_b200_dropin.so: the C++ host layer (namespace paris::b200 behind the
reference's backend contract + the stage wrappers + the reference-shaped per-projection loop)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import capi

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libparis_b200_dropin.so")
_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        capi.lib()  # the product library must be loaded (and exist) first
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `make lib`")
        L = C.CDLL(LIB_PATH)
        L.paris_b200_dropin_reconstruct.restype = C.c_int
        L.paris_b200_dropin_reconstruct.argtypes = [
            C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(capi.DetectorGeometry),
            C.POINTER(capi.VolumeGeometry), C.c_int, C.POINTER(capi.Roi), C.c_uint32, C.c_uint32, C.c_uint32,
            C.c_uint32, C.c_uint32, C.c_int, C.c_void_p, C.c_char_p, C.c_size_t]
        L.paris_b200_dropin_set_device.restype = C.c_int
        L.paris_b200_dropin_set_device.argtypes = [C.c_int]
        L.paris_b200_dropin_context.restype = C.c_void_p
        L.paris_b200_dropin_context.argtypes = []
        _lib = L
    return _lib


def set_device(device: int) -> None:
    """paris::b200::set_device for the calling thread (creates / rebinds its context)."""
    if lib().paris_b200_dropin_set_device(device) != 0:
        raise capi.Error(capi.ECUDA, capi.lib().paris_b200_last_error().decode())


def context_handle() -> int:
    """The paris_b200_ctx* of the calling thread inside the C++ layer."""
    h = lib().paris_b200_dropin_context()
    if not h:
        raise capi.Error(capi.ECUDA, capi.lib().paris_b200_last_error().decode())
    return h


def reconstruct(h_stack, n_proj: int, det: capi.DetectorGeometry, vol_full: capi.VolumeGeometry, h_region,
                region_dims, roi: capi.Roi | None = None, slab_id: int = 0, num_slabs: int = 1, device: int = 0,
                first_idx: int = 0, idx_stride: int = 1) -> None:
    """Run the reference-shaped loop (load -> weight -> filter -> backproject per projection, then
    copy_d2h) for one slab.  h_stack / h_region: numpy arrays (ideally pinned) or raw host addresses."""
    err = C.create_string_buffer(512)
    rc = lib().paris_b200_dropin_reconstruct(
        capi._ptr(h_stack), n_proj, first_idx, idx_stride, C.byref(det), C.byref(vol_full), int(roi is not None),
        C.byref(roi) if roi is not None else None, region_dims[0], region_dims[1], region_dims[2], slab_id, num_slabs,
        device, capi._ptr(h_region), err, len(err))
    if rc != 0:
        raise capi.Error(capi.ECUDA, err.value.decode())
