"""oracle -- CPU checkers for the FDK hot path.  TEST INFRASTRUCTURE, not product code.

Two libraries, both built by ``oracle/Makefile``:

* ``liboracle.so`` -- ``fdk_oracle.c``: a plain-C restatement of the reference's
  weight -> filter -> backproject path (every function cites the reference file:line).
* ``_ref/libparis_ref.so`` -- the reference's OWN OpenMP backend and wrappers, compiled
  unmodified from ``/root/reference/src`` (only where that tree exists; the built file
  travels to the GPU box).  It pins the restatement.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import this package.  ``paris_b200`` never does.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PORT_PATH = os.path.join(_HERE, "liboracle.so")
REF_PATH = os.path.join(_HERE, "_ref", "libparis_ref.so")


class DetectorGeometry(C.Structure):  # /root/reference/src/geometry.h:30-47
    _fields_ = [("n_row", C.c_uint32), ("n_col", C.c_uint32),
                ("l_px_row", C.c_float), ("l_px_col", C.c_float),
                ("delta_s", C.c_float), ("delta_t", C.c_float),
                ("d_so", C.c_float), ("d_od", C.c_float), ("delta_phi", C.c_float)]


class VolumeGeometry(C.Structure):  # /root/reference/src/geometry.h:49-58
    _fields_ = [("dim_x", C.c_uint32), ("dim_y", C.c_uint32), ("dim_z", C.c_uint32),
                ("l_vx_x", C.c_float), ("l_vx_y", C.c_float), ("l_vx_z", C.c_float)]


class Roi(C.Structure):  # /root/reference/src/region_of_interest.h:30-38
    _fields_ = [("x1", C.c_uint32), ("x2", C.c_uint32), ("y1", C.c_uint32),
                ("y2", C.c_uint32), ("z1", C.c_uint32), ("z2", C.c_uint32)]


class SubvolumeInfo(C.Structure):
    _fields_ = [("dim_x", C.c_uint32), ("dim_y", C.c_uint32), ("dim_z", C.c_uint32),
                ("remainder", C.c_uint32), ("num", C.c_int)]


def build(force: bool = False) -> None:
    """Compile liboracle.so (always possible) and _ref/libparis_ref.so (only where /root/reference exists)."""
    if force or not os.path.exists(PORT_PATH) or (os.path.isdir("/root/reference/src") and not os.path.exists(REF_PATH)):
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True,
                       stdout=subprocess.DEVNULL)


def have_ref() -> bool:
    return os.path.exists(REF_PATH)


def _fp(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_float))


def det_geo(n_row, n_col, l_px_row, l_px_col, delta_s, delta_t, d_so, d_od, delta_phi) -> DetectorGeometry:
    return DetectorGeometry(n_row, n_col, l_px_row, l_px_col, delta_s, delta_t, d_so, d_od, delta_phi)


class _Lib:
    """Common ctypes surface of the port (prefix ``oracle_``) and the reference build (``paris_ref_``)."""

    prefix = ""

    def __init__(self, path: str):
        self.lib = C.CDLL(path)
        p = self.prefix
        L = self.lib
        fp = C.POINTER(C.c_float)
        getattr(L, p + "calculate_volume_geometry").argtypes = [C.POINTER(DetectorGeometry), C.POINTER(VolumeGeometry)]
        getattr(L, p + "apply_roi").argtypes = [C.POINTER(VolumeGeometry), C.POINTER(Roi), C.POINTER(VolumeGeometry)]
        getattr(L, p + "weight").argtypes = [fp, C.POINTER(DetectorGeometry)]
        getattr(L, p + "filter").argtypes = [fp, C.POINTER(DetectorGeometry)]
        getattr(L, p + "make_filter").argtypes = [C.c_uint32, C.c_float, fp]
        getattr(L, p + "backproject").argtypes = [fp, C.c_uint32, C.c_float, C.c_int, fp, C.c_uint32, C.c_uint32,
                                                  C.c_uint32, C.c_uint32, C.POINTER(DetectorGeometry),
                                                  C.POINTER(VolumeGeometry), C.c_int, C.POINTER(Roi)]
        getattr(L, p + "reconstruct").argtypes = [fp, C.c_uint32, C.c_uint32, C.c_uint32, fp, C.c_uint32, C.c_uint32,
                                                  C.c_uint32, C.POINTER(DetectorGeometry), C.POINTER(VolumeGeometry),
                                                  C.c_int, C.POINTER(Roi), C.POINTER(C.c_double)]
        getattr(L, p + "num_threads").restype = C.c_int
        for name in ("calculate_volume_geometry", "apply_roi", "weight", "filter", "make_filter", "backproject",
                     "reconstruct"):
            getattr(L, p + name).restype = None

    def _f(self, name):
        return getattr(self.lib, self.prefix + name)

    def num_threads(self) -> int:
        return int(self._f("num_threads")())

    def set_num_threads(self, n: int) -> None:
        """OpenMP threads of THIS library's parallel regions (torchrun exports OMP_NUM_THREADS=1)."""
        f = self._f("set_num_threads")
        f.argtypes, f.restype = [C.c_int], None
        f(int(n))

    def calculate_volume_geometry(self, det: DetectorGeometry) -> VolumeGeometry:
        out = VolumeGeometry()
        self._f("calculate_volume_geometry")(C.byref(det), C.byref(out))
        return out

    def apply_roi(self, vol: VolumeGeometry, roi: Roi) -> VolumeGeometry:
        out = VolumeGeometry()
        self._f("apply_roi")(C.byref(vol), C.byref(roi), C.byref(out))
        return out

    def weight(self, proj: np.ndarray, det: DetectorGeometry) -> np.ndarray:
        """proj: (n_col, n_row) float32 (row index = detector row t, fastest index = s); returns a new array."""
        out = np.ascontiguousarray(proj, dtype=np.float32).copy()
        self._f("weight")(_fp(out), C.byref(det))
        return out

    def filter(self, proj: np.ndarray, det: DetectorGeometry) -> np.ndarray:
        out = np.ascontiguousarray(proj, dtype=np.float32).copy()
        self._f("filter")(_fp(out), C.byref(det))
        return out

    def make_filter(self, size: int, tau: float) -> np.ndarray:
        k = np.zeros(size // 2 + 1, dtype=np.float32)
        self._f("make_filter")(size, tau, _fp(k))
        return k

    def backproject(self, proj: np.ndarray, idx: int, vol: np.ndarray, det: DetectorGeometry,
                    vol_full: VolumeGeometry, roi: Roi | None = None, v_offset: int = 0,
                    phi_deg: float | None = None) -> None:
        """Accumulate one (filtered) projection into vol, shaped (dim_z, dim_y, dim_x) float32, in place."""
        assert vol.dtype == np.float32 and vol.flags["C_CONTIGUOUS"] and vol.ndim == 3
        r = roi if roi is not None else Roi()
        proj = np.ascontiguousarray(proj, dtype=np.float32)
        self._f("backproject")(_fp(proj), idx, 0.0 if phi_deg is None else phi_deg, int(phi_deg is not None),
                               _fp(vol), vol.shape[2], vol.shape[1], vol.shape[0], v_offset,
                               C.byref(det), C.byref(vol_full), int(roi is not None), C.byref(r))

    def reconstruct(self, stack: np.ndarray, vol_shape, det: DetectorGeometry, vol_full: VolumeGeometry,
                    roi: Roi | None = None, first_idx: int = 0, idx_stride: int = 1):
        """weight -> filter -> backproject over a raw stack (n_proj, n_col, n_row).
        Returns (volume (dz, dy, dx), [t_weight, t_filter, t_backproject] seconds, first projection excluded)."""
        assert stack.dtype == np.float32 and stack.flags["C_CONTIGUOUS"] and stack.ndim == 3
        vol = np.zeros(vol_shape, dtype=np.float32)
        r = roi if roi is not None else Roi()
        times = (C.c_double * 3)()
        self._f("reconstruct")(_fp(stack), stack.shape[0], first_idx, idx_stride, _fp(vol),
                               vol.shape[2], vol.shape[1], vol.shape[0], C.byref(det), C.byref(vol_full),
                               int(roi is not None), C.byref(r), times)
        return vol, [times[0], times[1], times[2]]


class Port(_Lib):
    """The plain-C restatement (oracle/fdk_oracle.c).  Stateless: any number of geometries per process."""

    prefix = "oracle_"

    def __init__(self):
        build()
        super().__init__(PORT_PATH)
        self.lib.oracle_filter_size.argtypes = [C.c_uint32]
        self.lib.oracle_filter_size.restype = C.c_uint32
        self.lib.oracle_apply_filter.argtypes = [C.POINTER(C.c_float), C.c_uint32, C.c_uint32,
                                                 C.POINTER(C.c_float), C.c_uint32]
        self.lib.oracle_apply_filter.restype = None
        self.lib.oracle_make_subvolume_information.argtypes = [C.POINTER(VolumeGeometry), C.c_int,
                                                               C.POINTER(SubvolumeInfo)]
        self.lib.oracle_make_subvolume_information.restype = None

    def filter_size(self, n_row: int) -> int:
        return int(self.lib.oracle_filter_size(n_row))

    def weight_filter_rows(self, proj: np.ndarray, det: DetectorGeometry, row0: int, n_rows: int) -> np.ndarray:
        """weight -> filter of rows [row0, row0 + n_rows) of a full-size projection; returns a copy whose other rows
        are untouched (bit-identical to the same rows of filter(weight(proj)))."""
        out = np.ascontiguousarray(proj, dtype=np.float32).copy()
        size = self.filter_size(det.n_row)
        k = self.make_filter(size, det.l_px_row)
        f = self.lib.oracle_weight_filter_rows
        f.argtypes = [C.POINTER(C.c_float), C.POINTER(DetectorGeometry), C.POINTER(C.c_float), C.c_uint32, C.c_uint32,
                      C.c_uint32]
        f.restype = None
        f(_fp(out), C.byref(det), _fp(k), size, row0, n_rows)
        return out

    def reconstruct_block(self, band_stack: np.ndarray, row0: int, det: DetectorGeometry, vol_full: VolumeGeometry,
                          roi: Roi, first_idx: int = 0, idx_stride: int = 1, vol: np.ndarray | None = None) -> np.ndarray:
        """The reference's loop for the box `roi` (ROI path) from the band of detector rows [row0, row0 + n_rows) of
        every raw projection: band_stack is (n_proj, n_rows, n_row).  Returns the box (dz, dy, dx); pass the previous
        result as `vol` to add the next chunk of projections (first_idx = scan index of its first projection)."""
        assert band_stack.dtype == np.float32 and band_stack.flags["C_CONTIGUOUS"] and band_stack.ndim == 3
        assert band_stack.shape[2] == det.n_row and row0 + band_stack.shape[1] <= det.n_col
        box = self.apply_roi(vol_full, roi)
        if vol is None:
            vol = np.zeros((box.dim_z, box.dim_y, box.dim_x), dtype=np.float32)
        assert vol.shape == (box.dim_z, box.dim_y, box.dim_x) and vol.dtype == np.float32 and vol.flags["C_CONTIGUOUS"]
        f = self.lib.oracle_reconstruct_block
        f.argtypes = [C.POINTER(C.c_float), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                      C.POINTER(C.c_float), C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(DetectorGeometry),
                      C.POINTER(VolumeGeometry), C.POINTER(Roi)]
        f.restype = None
        f(_fp(band_stack), band_stack.shape[0], first_idx, idx_stride, row0, band_stack.shape[1], _fp(vol),
          box.dim_x, box.dim_y, box.dim_z, C.byref(det), C.byref(vol_full), C.byref(roi))
        return vol

    def make_subvolume_information(self, vol: VolumeGeometry, num: int) -> SubvolumeInfo:
        out = SubvolumeInfo()
        self.lib.oracle_make_subvolume_information(C.byref(vol), num, C.byref(out))
        return out


class Reference(_Lib):
    """The reference's own compiled code (oracle/_ref/libparis_ref.so).

    The reference freezes geometry in function-local statics at first use, so every
    instance loads a PRIVATE COPY of the shared object: one instance == one geometry."""

    prefix = "paris_ref_"

    def __init__(self):
        build()
        if not have_ref():
            raise FileNotFoundError(REF_PATH)
        self._tmp = tempfile.NamedTemporaryFile(prefix="libparis_ref_", suffix=".so", delete=False)
        with open(REF_PATH, "rb") as src:
            shutil.copyfileobj(src, self._tmp)
        self._tmp.close()
        super().__init__(self._tmp.name)
        os.unlink(self._tmp.name)  # the mapping stays valid
