"""ctypes binding of the C ABI in include/paris_b200.h (libparis_b200.so).

This is plumbing for tests/ and bench.py: every call below is a 1:1 forward to the C entry point of
the same name.  There is no Python or CPU implementation behind it -- importing works without a GPU
(so the symbol table can be checked), every compute call fails loudly without an sm_100a device.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

# One hardware queue per stream (the default is 8): a group member keeps five streams busy next to whatever the host
# program runs, and a stream blocked on a peer's arrival flag must never hold back another stream that shares its queue.
# Read by the driver when the process creates its first CUDA context, hence set at import.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
# Kernels are loaded with their module, not lazily at first launch: a lazy load synchronises the context, which never
# returns inside a group step (streams waiting for flags that work not yet enqueued will set).  The library sets both
# defaults itself when it is loaded; here as well because torch may initialise CUDA before that.
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PARIS_B200_LIB") or os.path.join(_HERE, "libparis_b200.so")   # (override: kernel A/B experiments)

OK, EINVAL, ECUDA, ENOMEM, ESTATE = 0, 1, 2, 3, 4
BP_FUSE_WEIGHT_FILTER = 1


class DetectorGeometry(C.Structure):
    _fields_ = [("n_row", C.c_uint32), ("n_col", C.c_uint32),
                ("l_px_row", C.c_float), ("l_px_col", C.c_float),
                ("delta_s", C.c_float), ("delta_t", C.c_float),
                ("d_so", C.c_float), ("d_od", C.c_float), ("delta_phi", C.c_float)]


class VolumeGeometry(C.Structure):
    _fields_ = [("dim_x", C.c_uint32), ("dim_y", C.c_uint32), ("dim_z", C.c_uint32),
                ("l_vx_x", C.c_float), ("l_vx_y", C.c_float), ("l_vx_z", C.c_float)]


class Roi(C.Structure):
    _fields_ = [("x1", C.c_uint32), ("x2", C.c_uint32), ("y1", C.c_uint32),
                ("y2", C.c_uint32), ("z1", C.c_uint32), ("z2", C.c_uint32)]


class SubvolumeInfo(C.Structure):
    _fields_ = [("dim_x", C.c_uint32), ("dim_y", C.c_uint32), ("dim_z", C.c_uint32),
                ("remainder", C.c_uint32), ("num", C.c_int32)]


GROUP_HANDLE_BYTES, GROUP_MAX_MEMBERS, GROUP_MAX_ROUNDS = 256, 64, 256
SAMPLES_F32, SAMPLES_U16 = 0, 1
EXCHANGE_COPY_ENGINE, EXCHANGE_KERNEL = 0, 1


class GroupConfig(C.Structure):
    _fields_ = [("rank", C.c_int32), ("world", C.c_int32), ("det", DetectorGeometry), ("vol_full", VolumeGeometry),
                ("enable_roi", C.c_int32), ("roi", Roi), ("n_proj", C.c_uint32), ("angles_deg", C.POINTER(C.c_float)),
                ("slabs_per_member", C.c_uint32), ("stream_slabs", C.c_uint32), ("sample_type", C.c_uint32),
                ("first_round", C.c_uint32), ("max_round", C.c_uint32), ("whole_projections", C.c_uint32),
                ("exchange", C.c_uint32), ("x_parts", C.c_uint32), ("host_row_floats", C.c_uint32)]


class GroupPlan(C.Structure):
    _fields_ = [("region_x", C.c_uint32), ("region_y", C.c_uint32), ("region_z", C.c_uint32), ("region_z0", C.c_uint32),
                ("layout", C.c_uint32), ("pitch", C.c_uint32), ("slabs_total", C.c_uint32), ("slab_dz", C.c_uint32),
                ("slab_remainder", C.c_uint32), ("x_parts", C.c_uint32), ("x_dx", C.c_uint32), ("x_remainder", C.c_uint32),
                ("rounds", C.c_uint32),
                ("round_first", C.c_uint32 * GROUP_MAX_ROUNDS), ("round_count", C.c_uint32 * GROUP_MAX_ROUNDS),
                ("band_lo", C.c_uint32 * GROUP_MAX_MEMBERS), ("band_hi", C.c_uint32 * GROUP_MAX_MEMBERS)]


class GroupInfo(C.Structure):
    _fields_ = [("my_projections", C.c_uint32), ("rounds", C.c_uint32), ("slabs", C.c_uint32), ("z_first", C.c_uint32),
                ("z_count", C.c_uint32), ("x_first", C.c_uint32), ("x_count", C.c_uint32), ("region_x", C.c_uint32), ("region_y", C.c_uint32), ("region_z", C.c_uint32),
                ("band_lo", C.c_uint32), ("band_hi", C.c_uint32), ("layout", C.c_uint32), ("pitch", C.c_uint32),
                ("d_stack", C.c_void_p), ("slab_buffers", C.c_uint32), ("d_first_slab", C.c_void_p),
                ("bytes_pushed", C.c_uint64), ("ctx", C.c_void_p), ("filter_ctx", C.c_void_p), ("memops", C.c_uint32)]


class Error(RuntimeError):
    def __init__(self, code: int, text: str):
        super().__init__(f"paris_b200 error {code}: {text}")
        self.code = code


_P = C.POINTER
_vp = C.c_void_p
_fp = C.c_void_p      # device / host float pointers travel as integers
_u32 = C.c_uint32
_f = C.c_float

# name -> (restype, argtypes); must list EVERY symbol include/paris_b200.h declares (tests check that)
SIGNATURES = {
    "paris_b200_last_error": (C.c_char_p, []),
    "paris_b200_version": (C.c_char_p, []),
    "paris_b200_device_count": (C.c_int, [_P(C.c_int)]),
    "paris_b200_ctx_create": (C.c_int, [C.c_int, _P(_vp)]),
    "paris_b200_ctx_destroy": (C.c_int, [_vp]),
    "paris_b200_ctx_device": (C.c_int, [_vp, _P(C.c_int)]),
    "paris_b200_ctx_bind": (C.c_int, [_vp]),
    "paris_b200_ctx_sync": (C.c_int, [_vp]),
    "paris_b200_ctx_stream": (C.c_int, [_vp, _P(_vp)]),
    "paris_b200_ctx_launch_count": (C.c_int, [_vp, _P(C.c_uint64)]),
    "paris_b200_ctx_stats": (C.c_int, [_vp, _P(C.c_uint64), C.c_int]),
    "paris_b200_ctx_bp_kernel_info": (C.c_int, [_vp, C.c_char_p, C.c_size_t, _P(C.c_uint64), _P(C.c_uint64)]),
    "paris_b200_event_create": (C.c_int, [_vp, _P(_vp)]),
    "paris_b200_event_record": (C.c_int, [_vp, _vp]),
    "paris_b200_event_elapsed_ms": (C.c_int, [_vp, _vp, _P(_f)]),
    "paris_b200_event_destroy": (C.c_int, [_vp]),
    "paris_b200_ctx_set_option": (C.c_int, [_vp, C.c_char_p, C.c_int64]),
    "paris_b200_calculate_volume_geometry": (C.c_int, [_P(DetectorGeometry), _P(VolumeGeometry)]),
    "paris_b200_apply_roi": (C.c_int, [_P(VolumeGeometry), _P(Roi), _P(VolumeGeometry)]),
    "paris_b200_filter_size": (_u32, [_u32]),
    "paris_b200_make_subvolume_information": (C.c_int, [_vp, _P(VolumeGeometry), _P(DetectorGeometry), C.c_int,
                                                        _P(SubvolumeInfo)]),
    "paris_b200_host_alloc": (C.c_int, [C.c_size_t, C.c_int, _P(_vp)]),
    "paris_b200_host_free": (C.c_int, [_vp]),
    "paris_b200_host_register": (C.c_int, [_vp, C.c_size_t]),
    "paris_b200_host_unregister": (C.c_int, [_vp]),
    "paris_b200_dev_alloc": (C.c_int, [_vp, C.c_size_t, _P(_vp)]),
    "paris_b200_dev_free": (C.c_int, [_vp, _vp]),
    "paris_b200_volume_alloc": (C.c_int, [_vp, _u32, _u32, _u32, _P(_vp)]),
    "paris_b200_volume_free": (C.c_int, [_vp, _fp]),
    "paris_b200_volume_clear": (C.c_int, [_vp, _fp, _u32, _u32, _u32]),
    "paris_b200_proj_h2d": (C.c_int, [_vp, _fp, _fp, _u32, _u32]),
    "paris_b200_h2d_done": (C.c_int, [_vp, _P(C.c_int)]),
    "paris_b200_proj_d2h": (C.c_int, [_vp, _fp, _fp, _u32, _u32]),
    "paris_b200_vol_h2d": (C.c_int, [_vp, _fp, _fp, C.c_size_t]),
    "paris_b200_vol_d2h": (C.c_int, [_vp, _fp, _fp, C.c_size_t]),
    "paris_b200_weight": (C.c_int, [_vp, _fp, _u32, _u32, _f, _f, _f, _f, _f]),
    "paris_b200_filter_create": (C.c_int, [_vp, _u32, _f, _P(_vp)]),
    "paris_b200_filter_destroy": (C.c_int, [_vp]),
    "paris_b200_filter_read": (C.c_int, [_vp, _vp, _fp]),
    "paris_b200_apply_filter": (C.c_int, [_vp, _fp, _u32, _u32, _vp, _u32, _u32]),
    "paris_b200_weight_filter": (C.c_int, [_vp, _fp, _u32, _u32, _f, _f, _f, _f, _f, _vp, _u32]),
    "paris_b200_backproject": (C.c_int, [_vp, _fp, _u32, _u32, _fp, _u32, _u32, _u32, _u32, _P(DetectorGeometry),
                                         _P(VolumeGeometry), C.c_int, _P(Roi), _f, _f, _f, _f, _u32, _vp, _vp]),
    "paris_b200_h2d_wait": (C.c_int, [_vp]),
    "paris_b200_flush": (C.c_int, [_vp]),
    "paris_b200_stack_slot_bytes": (C.c_int, [_u32, _u32, _P(C.c_size_t), _P(_u32)]),
    "paris_b200_choose_stack_layout": (C.c_int, [_P(DetectorGeometry), _P(VolumeGeometry), _P(_u32)]),
    "paris_b200_stack_alloc": (C.c_int, [_vp, _u32, _u32, _u32, _P(_vp)]),
    "paris_b200_stack_free": (C.c_int, [_vp, _fp]),
    "paris_b200_filter_to_stack": (C.c_int, [_vp, _fp, _P(DetectorGeometry), _vp, _fp, _u32, _u32]),
    "paris_b200_filter_to_stack_batch": (C.c_int, [_vp, _fp, C.c_size_t, _u32, _P(DetectorGeometry), _vp, _fp, _u32,
                                                   _u32]),
    "paris_b200_filter_to_stack_batch_u16": (C.c_int, [_vp, _vp, C.c_size_t, _u32, _P(DetectorGeometry), _vp, _fp, _u32, _u32]),
    "paris_b200_backproject_stack": (C.c_int, [_vp, _fp, _u32, _u32, _P(_f), _P(_f), _fp, _u32, _u32, _u32, _u32,
                                               _P(DetectorGeometry), _P(VolumeGeometry), C.c_int, _P(Roi), _u32]),
    "paris_b200_backproject_stack_d2h": (C.c_int, [_vp, _fp, _u32, _u32, _P(_f), _P(_f), _fp, _u32, _u32, _u32, _u32,
                                                   _P(DetectorGeometry), _P(VolumeGeometry), C.c_int, _P(Roi), _u32,
                                                   _fp]),
    "paris_b200_group_plan": (C.c_int, [_P(GroupConfig), _P(GroupPlan)]),
    "paris_b200_group_share": (C.c_int, [_P(GroupPlan), _u32, _u32, _u32, _P(_u32), _P(_u32)]),
    "paris_b200_group_handle_bytes": (C.c_size_t, []),
    "paris_b200_group_create": (C.c_int, [C.c_int, _P(GroupConfig), _P(_vp)]),
    "paris_b200_group_destroy": (C.c_int, [_vp]),
    "paris_b200_group_export": (C.c_int, [_vp, C.c_char_p, C.c_size_t]),
    "paris_b200_group_connect": (C.c_int, [_vp, C.c_char_p, C.c_size_t]),
    "paris_b200_group_info": (C.c_int, [_vp, _P(GroupInfo)]),
    "paris_b200_group_set_host_row": (C.c_int, [_vp, _u32]),
    "paris_b200_group_projection_index": (C.c_int, [_vp, _u32, _P(_u32)]),
    "paris_b200_group_begin": (C.c_int, [_vp, _P(_vp), _vp, _vp]),
    "paris_b200_group_step_open": (C.c_int, [_vp, _vp]),
    "paris_b200_group_step_round": (C.c_int, [_vp, _u32, _P(_vp), _vp]),
    "paris_b200_group_uploaded": (C.c_int, [_vp, _u32, _P(C.c_int)]),
    "paris_b200_group_step_finish": (C.c_int, [_vp]),
    "paris_b200_group_end": (C.c_int, [_vp]),
    "paris_b200_group_reconstruct": (C.c_int, [_vp, _P(_vp), _vp, _vp]),
    "paris_b200_group_debug_state": (C.c_int, [_vp, _P(_u32), _u32]),
    "paris_b200_phantom_project": (C.c_int, [_vp, _P(C.c_double), _u32, _P(DetectorGeometry), _u32, _u32, _fp]),
}

_lib = None


def lib() -> C.CDLL:
    """Load libparis_b200.so (built by `make lib` / __graft_entry__.build()).  No fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `make lib` (or __graft_entry__.build()); "
                              "paris_b200 has no CPU fallback")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != OK:
        raise Error(rc, lib().paris_b200_last_error().decode())


def _ptr(a) -> int:
    """Address of a numpy array's data, or pass an integer device pointer through."""
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data
    return int(a)


def device_count() -> int:
    n = C.c_int(0)
    check(lib().paris_b200_device_count(C.byref(n)))
    return n.value


def filter_size(n_row: int) -> int:
    return int(lib().paris_b200_filter_size(n_row))


def calculate_volume_geometry(det: DetectorGeometry) -> VolumeGeometry:
    out = VolumeGeometry()
    check(lib().paris_b200_calculate_volume_geometry(C.byref(det), C.byref(out)))
    return out


def apply_roi(vol: VolumeGeometry, roi: Roi) -> VolumeGeometry:
    out = VolumeGeometry()
    check(lib().paris_b200_apply_roi(C.byref(vol), C.byref(roi), C.byref(out)))
    return out


LAYOUT_PLAIN, LAYOUT_SPLIT2 = 0, 1


def choose_stack_layout(det: DetectorGeometry, vol_full: VolumeGeometry) -> int:
    out = _u32(0)
    check(lib().paris_b200_choose_stack_layout(C.byref(det), C.byref(vol_full), C.byref(out)))
    return out.value


def stack_slot_bytes(n_row: int, n_col: int):
    b, p = C.c_size_t(0), _u32(0)
    check(lib().paris_b200_stack_slot_bytes(n_row, n_col, C.byref(b), C.byref(p)))
    return b.value, p.value


class PinnedArray:
    """A float32 (or uint16) numpy view over pinned host memory from paris_b200_host_alloc."""

    def __init__(self, shape, zero=False, dtype=np.float32):
        n = int(np.prod(shape))
        p = _vp()
        ctype = {np.dtype(np.float32): C.c_float, np.dtype(np.uint16): C.c_uint16}[np.dtype(dtype)]
        check(lib().paris_b200_host_alloc(n * C.sizeof(ctype), int(zero), C.byref(p)))
        self.ptr = p.value
        self.array = np.ctypeslib.as_array((ctype * n).from_address(self.ptr)).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            check(lib().paris_b200_host_free(self.ptr))
            self.ptr = 0


class Context:
    """One per device; thin object wrapper over the paris_b200_ctx_* / stage entry points."""

    def __init__(self, device: int = 0, handle: int | None = None):
        """handle: wrap an existing paris_b200_ctx* (not owned, e.g. the C++ layer's thread context)."""
        self._L = lib()
        self._owned = handle is None
        if handle is None:
            h = _vp()
            check(self._L.paris_b200_ctx_create(device, C.byref(h)))
            self.h = h
        else:
            self.h = _vp(handle)
        self.device = device

    def close(self):
        if self.h and self._owned:
            check(self._L.paris_b200_ctx_destroy(self.h))
        self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- plumbing
    def sync(self):
        check(self._L.paris_b200_ctx_sync(self.h))

    def stream(self) -> int:
        s = _vp()
        check(self._L.paris_b200_ctx_stream(self.h, C.byref(s)))
        return s.value or 0

    def launch_count(self) -> int:
        n = C.c_uint64(0)
        check(self._L.paris_b200_ctx_launch_count(self.h, C.byref(n)))
        return n.value

    def stats(self) -> dict:
        a = (C.c_uint64 * 6)()
        check(self._L.paris_b200_ctx_stats(self.h, a, 6))
        return dict(zip(("launches", "pool_malloc", "pool_ready", "pool_busy", "flushes", "pool_size"), list(a)))

    def bp_kernel_info(self) -> dict:
        """Instantiation of the most recent backprojection launch, launches that went to the TMA / exact kernel."""
        name = C.create_string_buffer(96)
        tma, exact = C.c_uint64(0), C.c_uint64(0)
        check(self._L.paris_b200_ctx_bp_kernel_info(self.h, name, 96, C.byref(tma), C.byref(exact)))
        return {"last": name.value.decode(), "tma_launches": tma.value, "exact_launches": exact.value}

    def event(self) -> int:
        """Create a CUDA event and record it on the compute stream."""
        e = _vp()
        check(self._L.paris_b200_event_create(self.h, C.byref(e)))
        check(self._L.paris_b200_event_record(self.h, e))
        return e.value

    def elapsed_ms(self, start: int, stop: int, destroy: bool = True) -> float:
        ms = _f(0)
        check(self._L.paris_b200_event_elapsed_ms(start, stop, C.byref(ms)))
        if destroy:
            check(self._L.paris_b200_event_destroy(start))
            check(self._L.paris_b200_event_destroy(stop))
        return float(ms.value)

    def set_option(self, name: str, value: int):
        check(self._L.paris_b200_ctx_set_option(self.h, name.encode(), value))

    def make_subvolume_information(self, vol: VolumeGeometry, det: DetectorGeometry, num_slabs: int) -> SubvolumeInfo:
        out = SubvolumeInfo()
        check(self._L.paris_b200_make_subvolume_information(self.h, C.byref(vol), C.byref(det), num_slabs,
                                                            C.byref(out)))
        return out

    # -- memory
    def dev_alloc(self, nbytes: int) -> int:
        p = _vp()
        check(self._L.paris_b200_dev_alloc(self.h, nbytes, C.byref(p)))
        return p.value

    def dev_free(self, d_ptr: int):
        check(self._L.paris_b200_dev_free(self.h, d_ptr))

    def volume_alloc(self, dim_x: int, dim_y: int, dim_z: int) -> int:
        p = _vp()
        check(self._L.paris_b200_volume_alloc(self.h, dim_x, dim_y, dim_z, C.byref(p)))
        return p.value

    def volume_clear(self, d_vol: int, dim_x: int, dim_y: int, dim_z: int):
        check(self._L.paris_b200_volume_clear(self.h, d_vol, dim_x, dim_y, dim_z))

    def volume_free(self, d_vol: int):
        check(self._L.paris_b200_volume_free(self.h, d_vol))

    def proj_h2d(self, h_src, d_dst: int, dim_x: int, dim_y: int):
        check(self._L.paris_b200_proj_h2d(self.h, _ptr(h_src), d_dst, dim_x, dim_y))

    def h2d_done(self) -> bool:
        d = C.c_int(0)
        check(self._L.paris_b200_h2d_done(self.h, C.byref(d)))
        return bool(d.value)

    def proj_d2h(self, d_src: int, h_dst, dim_x: int, dim_y: int):
        check(self._L.paris_b200_proj_d2h(self.h, d_src, _ptr(h_dst), dim_x, dim_y))

    def vol_h2d(self, h_src, d_dst: int, n_voxels: int):
        check(self._L.paris_b200_vol_h2d(self.h, _ptr(h_src), d_dst, n_voxels))

    def vol_d2h(self, d_src: int, h_dst, n_voxels: int):
        check(self._L.paris_b200_vol_d2h(self.h, d_src, _ptr(h_dst), n_voxels))

    # -- stages
    def weight(self, d_proj: int, dim_x: int, dim_y: int, h_min, v_min, d_sd, l_px_row, l_px_col):
        check(self._L.paris_b200_weight(self.h, d_proj, dim_x, dim_y, h_min, v_min, d_sd, l_px_row, l_px_col))

    def filter_create(self, size: int, tau: float) -> int:
        f = _vp()
        check(self._L.paris_b200_filter_create(self.h, size, tau, C.byref(f)))
        return f.value

    def filter_destroy(self, f: int):
        check(self._L.paris_b200_filter_destroy(f))

    def filter_read(self, f: int, size: int) -> np.ndarray:
        k = np.zeros(size // 2 + 1, dtype=np.float32)
        check(self._L.paris_b200_filter_read(self.h, f, _ptr(k)))
        return k

    def apply_filter(self, d_proj: int, dim_x: int, dim_y: int, f: int, size: int, n_col: int):
        check(self._L.paris_b200_apply_filter(self.h, d_proj, dim_x, dim_y, f, size, n_col))

    def weight_filter(self, d_proj: int, dim_x: int, dim_y: int, h_min, v_min, d_sd, l_px_row, l_px_col, f: int,
                      size: int):
        check(self._L.paris_b200_weight_filter(self.h, d_proj, dim_x, dim_y, h_min, v_min, d_sd, l_px_row, l_px_col,
                                               f, size))

    def backproject(self, d_proj: int, dim_x: int, dim_y: int, d_vol: int, v_dims, v_offset: int,
                    det: DetectorGeometry, vol_full: VolumeGeometry, roi: Roi | None, sin_phi: float, cos_phi: float,
                    delta_s_mm: float, delta_t_mm: float, flags: int = 0, filt: int | None = None):
        check(self._L.paris_b200_backproject(self.h, d_proj, dim_x, dim_y, d_vol, v_dims[0], v_dims[1], v_dims[2],
                                             v_offset, C.byref(det), C.byref(vol_full), int(roi is not None),
                                             C.byref(roi) if roi is not None else None, sin_phi, cos_phi,
                                             delta_s_mm, delta_t_mm, flags, filt, None))

    def flush(self):
        check(self._L.paris_b200_flush(self.h))

    def stack_alloc(self, n_row: int, n_col: int, slots: int) -> int:
        p = _vp()
        check(self._L.paris_b200_stack_alloc(self.h, n_row, n_col, slots, C.byref(p)))
        return p.value

    def stack_free(self, d_stack: int):
        check(self._L.paris_b200_stack_free(self.h, d_stack))

    def filter_to_stack(self, d_raw: int, det: DetectorGeometry, filt: int, d_stack: int, slot: int,
                        layout: int = LAYOUT_PLAIN):
        check(self._L.paris_b200_filter_to_stack(self.h, d_raw, C.byref(det), filt, d_stack, slot, layout))

    def filter_to_stack_batch(self, d_raw: int, raw_stride: int, count: int, det: DetectorGeometry, filt: int,
                              d_stack: int, first_slot: int, layout: int = LAYOUT_PLAIN):
        check(self._L.paris_b200_filter_to_stack_batch(self.h, d_raw, raw_stride, count, C.byref(det), filt, d_stack,
                                                       first_slot, layout))

    def filter_to_stack_batch_u16(self, d_raw: int, raw_stride: int, count: int, det: DetectorGeometry, filt: int,
                                  d_stack: int, first_slot: int, layout: int = LAYOUT_PLAIN):
        """d_raw: uint16 samples, raw_stride in samples"""
        check(self._L.paris_b200_filter_to_stack_batch_u16(self.h, d_raw, raw_stride, count, C.byref(det), filt,
                                                           d_stack, first_slot, layout))

    def backproject_stack(self, d_stack: int, first: int, count: int, sin_phi: np.ndarray, cos_phi: np.ndarray,
                          d_vol: int, v_dims, v_offset: int, det: DetectorGeometry, vol_full: VolumeGeometry,
                          roi: Roi | None = None, layout: int = LAYOUT_PLAIN):
        sn = np.ascontiguousarray(sin_phi, dtype=np.float32)
        cs = np.ascontiguousarray(cos_phi, dtype=np.float32)
        assert sn.size >= count and cs.size >= count
        check(self._L.paris_b200_backproject_stack(self.h, d_stack, first, count,
                                                   sn.ctypes.data_as(_P(_f)), cs.ctypes.data_as(_P(_f)),
                                                   d_vol, v_dims[0], v_dims[1], v_dims[2], v_offset,
                                                   C.byref(det), C.byref(vol_full), int(roi is not None),
                                                   C.byref(roi) if roi is not None else None, layout))

    def backproject_stack_d2h(self, d_stack: int, first: int, count: int, sin_phi: np.ndarray, cos_phi: np.ndarray,
                              d_vol: int, v_dims, v_offset: int, det: DetectorGeometry, vol_full: VolumeGeometry,
                              h_dst: int, roi: Roi | None = None, layout: int = LAYOUT_PLAIN):
        """backproject_stack followed by the download to pinned host memory at h_dst, overlapped chunk by chunk."""
        sn = np.ascontiguousarray(sin_phi, dtype=np.float32)
        cs = np.ascontiguousarray(cos_phi, dtype=np.float32)
        assert sn.size >= count and cs.size >= count
        check(self._L.paris_b200_backproject_stack_d2h(self.h, d_stack, first, count,
                                                       sn.ctypes.data_as(_P(_f)), cs.ctypes.data_as(_P(_f)),
                                                       d_vol, v_dims[0], v_dims[1], v_dims[2], v_offset,
                                                       C.byref(det), C.byref(vol_full), int(roi is not None),
                                                       C.byref(roi) if roi is not None else None, layout, h_dst))

    def phantom_project(self, ellipsoids_mm: np.ndarray, det: DetectorGeometry, first_idx: int, n_proj: int,
                        d_stack_raw: int):
        e = np.ascontiguousarray(ellipsoids_mm, dtype=np.float64)
        check(self._L.paris_b200_phantom_project(self.h, e.ctypes.data_as(_P(C.c_double)), e.shape[0], C.byref(det),
                                                 first_idx, n_proj, d_stack_raw))


# ---- one scan across the GPUs of a box (csrc/group.cu) -----------------------------------------------------------

def group_config(rank: int, world: int, det: DetectorGeometry, vol_full: VolumeGeometry, n_proj: int, roi: Roi | None = None,
                 slabs_per_member: int = 1, stream_slabs: bool = False, first_round: int = 0, max_round: int = 0,
                 whole_projections: bool = False, exchange: int = EXCHANGE_COPY_ENGINE, angles_deg=None, x_parts: int = 1,
                 host_row_floats: int = 0, sample_type: int = SAMPLES_F32) -> GroupConfig:
    cfg = GroupConfig()
    cfg.rank, cfg.world, cfg.det, cfg.vol_full, cfg.n_proj = rank, world, det, vol_full, n_proj
    cfg.enable_roi = int(roi is not None)
    if roi is not None:
        cfg.roi = roi
    cfg.slabs_per_member, cfg.stream_slabs, cfg.sample_type = slabs_per_member, int(stream_slabs), sample_type
    cfg.first_round, cfg.max_round = first_round, max_round
    cfg.whole_projections, cfg.exchange = int(whole_projections), exchange
    cfg.x_parts, cfg.host_row_floats = x_parts, host_row_floats
    if angles_deg is not None:
        a = np.ascontiguousarray(angles_deg, dtype=np.float32)
        assert a.size == n_proj
        cfg._angles = a                     # (keeps the array alive as long as the struct)
        cfg.angles_deg = a.ctypes.data_as(C.POINTER(C.c_float))
    return cfg


def group_plan(cfg: GroupConfig) -> GroupPlan:
    """Rounds, slabs and bands of a configuration: host arithmetic only, works without a GPU."""
    plan = GroupPlan()
    check(lib().paris_b200_group_plan(C.byref(cfg), C.byref(plan)))
    return plan


def group_share(plan: GroupPlan, world: int, rnd: int, member: int):
    first, count = _u32(0), _u32(0)
    check(lib().paris_b200_group_share(C.byref(plan), world, rnd, member, C.byref(first), C.byref(count)))
    return first.value, count.value


class Group:
    """One member of a reconstruction group (paris_b200_group_*)."""

    def __init__(self, device: int, cfg: GroupConfig):
        self._L = lib()
        self.cfg = cfg
        h = _vp()
        check(self._L.paris_b200_group_create(device, C.byref(cfg), C.byref(h)))
        self.h = h
        self.device = device

    def export(self) -> bytes:
        buf = C.create_string_buffer(GROUP_HANDLE_BYTES)
        check(self._L.paris_b200_group_export(self.h, buf, GROUP_HANDLE_BYTES))
        return buf.raw

    def connect(self, handles):
        """handles: one export per member, in rank order"""
        blob = b"".join(handles)
        assert len(blob) == GROUP_HANDLE_BYTES * self.cfg.world
        check(self._L.paris_b200_group_connect(self.h, blob, GROUP_HANDLE_BYTES))

    def info(self) -> GroupInfo:
        out = GroupInfo()
        check(self._L.paris_b200_group_info(self.h, C.byref(out)))
        return out

    def set_host_row(self, host_row_floats: int):
        check(self._L.paris_b200_group_set_host_row(self.h, host_row_floats))

    def projection_index(self, local: int) -> int:
        out = _u32(0)
        check(self._L.paris_b200_group_projection_index(self.h, local, C.byref(out)))
        return out.value

    def _args(self, h_raw, d_raw, h_slabs):
        ptrs = None
        if h_raw is not None:
            ptrs = (_vp * len(h_raw))(*[int(p) for p in h_raw])
        return ptrs, (None if d_raw is None else int(d_raw)), (None if h_slabs is None else int(h_slabs))

    def begin(self, h_raw=None, d_raw=None, h_slabs=None):
        """h_raw: host addresses of this member's raw projections in local order (pinned), or d_raw: device address
        of the same, contiguous; h_slabs: host address for the member's slabs (pinned) or None."""
        ptrs, d, h = self._args(h_raw, d_raw, h_slabs)
        self._keep = ptrs
        check(self._L.paris_b200_group_begin(self.h, ptrs, d, h))

    def end(self):
        check(self._L.paris_b200_group_end(self.h))

    def reconstruct(self, h_raw=None, d_raw=None, h_slabs=None):
        self.begin(h_raw, d_raw, h_slabs)
        self.end()

    def debug_state(self) -> dict:
        n = 5 + 2 * self.cfg.world
        out = (_u32 * n)()
        check(self._L.paris_b200_group_debug_state(self.h, out, n))
        v = list(out)
        return {"busy": dict(zip(("backprojection", "download", "filter", "upload", "exchange"), v[:5])),
                "arrived": v[5:5 + self.cfg.world], "consumed": v[5 + self.cfg.world:]}

    def close(self):
        if self.h:
            check(self._L.paris_b200_group_destroy(self.h))
            self.h = None
