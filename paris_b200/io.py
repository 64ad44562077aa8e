"""ctypes binding of the I/O layer of the C++ host side (paris_b200/cpp/io, program_options) in
libparis_b200_dropin.so: HIS reader, DDBVF container, directory and angle helpers, option parser, projection
source.  The command-line driver itself is paris_b200/bin/paris_b200 (cpp/main.cpp)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import capi, dropin

CLI_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bin", "paris_b200")
_ready = False
_fp, _u32p = C.POINTER(C.c_float), C.POINTER(C.c_uint32)


class Options(C.Structure):
    _fields_ = [("det", capi.DetectorGeometry), ("enable_io", C.c_int), ("enable_roi", C.c_int),
                ("enable_angles", C.c_int), ("roi", capi.Roi), ("quality", C.c_uint32),
                ("input_path", C.c_char * 512), ("output_path", C.c_char * 512), ("prefix", C.c_char * 128),
                ("angle_path", C.c_char * 512)]


def _lib() -> C.CDLL:
    global _ready
    L = dropin.lib()
    if not _ready:
        err = [C.c_char_p, C.c_size_t]
        L.paris_b200_io_his_info.argtypes = [C.c_char_p, _u32p] + err
        L.paris_b200_io_his_read.argtypes = [C.c_char_p, _fp, C.c_uint32] + err
        L.paris_b200_io_ddbvf_create.restype = C.c_void_p
        L.paris_b200_io_ddbvf_create.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.c_uint32] + err
        L.paris_b200_io_ddbvf_open.restype = C.c_void_p
        L.paris_b200_io_ddbvf_open.argtypes = [C.c_char_p, _u32p] + err
        L.paris_b200_io_ddbvf_write.argtypes = [C.c_void_p, _fp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32] + err
        L.paris_b200_io_ddbvf_read.argtypes = [C.c_void_p, _fp, C.c_uint32, C.c_uint32] + err
        L.paris_b200_io_ddbvf_close.restype = None
        L.paris_b200_io_ddbvf_close.argtypes = [C.c_void_p]
        L.paris_b200_io_read_directory.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t] + err
        L.paris_b200_io_create_directory.argtypes = [C.c_char_p] + err
        L.paris_b200_io_read_angles.argtypes = [C.c_char_p, _fp, C.c_uint32] + err
        L.paris_b200_io_parse_options.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.POINTER(Options), C.c_char_p,
                                                  C.c_size_t]
        L.paris_b200_io_source_walk.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_uint32, _u32p, _fp,
                                                C.POINTER(C.c_int), _fp, C.c_uint32] + err
        L.paris_b200_io_scan_index.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_uint32, _u32p, _fp,
                                               C.POINTER(C.c_int), C.c_uint32, _u32p] + err
        L.paris_b200_io_scan_frame.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, _fp, C.POINTER(C.c_uint16)] + err
        _ready = True
    return L


class IoError(RuntimeError):
    pass


def _err():
    return C.create_string_buffer(1024)


def his_info(path: str):
    """(width, height, frames, number_type) or None for a file that is not a supported HIS file."""
    out, e = (C.c_uint32 * 4)(), _err()
    rc = _lib().paris_b200_io_his_info(path.encode(), out, e, len(e))
    if rc < 0:
        raise IoError(e.value.decode())
    return tuple(out) if rc == 1 else None


def his_read(path: str) -> np.ndarray:
    """(frames, height, width) float32; empty for an invalid file; IoError if it cannot be opened."""
    info = his_info(path)
    if info is None:
        return np.zeros((0, 0, 0), np.float32)
    w, h, n, _ = info
    out, e = np.zeros((n, h, w), np.float32), _err()
    got = _lib().paris_b200_io_his_read(path.encode(), out.ctypes.data_as(_fp), n, e, len(e))
    if got < 0:
        raise IoError(e.value.decode())
    return out[:got]


class Ddbvf:
    def __init__(self, handle, dims):
        self.h, self.dims = handle, tuple(dims)

    @classmethod
    def create(cls, path_prefix: str, dim_x: int, dim_y: int, dim_z: int) -> "Ddbvf":
        e = _err()
        h = _lib().paris_b200_io_ddbvf_create(path_prefix.encode(), dim_x, dim_y, dim_z, e, len(e))
        if not h:
            raise IoError(e.value.decode())
        return cls(h, (dim_x, dim_y, dim_z))

    @classmethod
    def open(cls, path: str) -> "Ddbvf":
        dims, e = (C.c_uint32 * 3)(), _err()
        h = _lib().paris_b200_io_ddbvf_open(path.encode(), dims, e, len(e))
        if not h:
            raise IoError(e.value.decode())
        return cls(h, dims)

    def write(self, vol: np.ndarray, first: int) -> None:
        vol = np.ascontiguousarray(vol, np.float32)
        e = _err()
        if _lib().paris_b200_io_ddbvf_write(self.h, vol.ctypes.data_as(_fp), vol.shape[2], vol.shape[1], vol.shape[0],
                                            first, e, len(e)) != 0:
            raise IoError(e.value.decode())

    def read(self, first: int, count: int) -> np.ndarray:
        out, e = np.zeros((count, self.dims[1], self.dims[0]), np.float32), _err()
        if _lib().paris_b200_io_ddbvf_read(self.h, out.ctypes.data_as(_fp), first, count, e, len(e)) != 0:
            raise IoError(e.value.decode())
        return out

    def close(self) -> None:
        if self.h:
            _lib().paris_b200_io_ddbvf_close(self.h)
            self.h = None


def read_directory(path: str) -> list[str]:
    out, e = C.create_string_buffer(1 << 16), _err()
    n = _lib().paris_b200_io_read_directory(path.encode(), out, len(out), e, len(e))
    if n < 0:
        raise IoError(e.value.decode())
    return [p for p in out.value.decode().split("\n") if p]


def create_directory(path: str) -> bool:
    e = _err()
    rc = _lib().paris_b200_io_create_directory(path.encode(), e, len(e))
    if rc < 0:
        raise IoError(e.value.decode())
    return rc == 1


def read_angles(path: str, capacity: int = 1 << 16) -> np.ndarray:
    out, e = np.zeros(capacity, np.float32), _err()
    n = _lib().paris_b200_io_read_angles(path.encode(), out.ctypes.data_as(_fp), capacity, e, len(e))
    if n < 0:
        raise IoError(e.value.decode())
    return out[:n].copy()


def parse_options(args: list[str]):
    """(result, Options, message): result 0 = ok, 1 = the program would print help and exit(0), 2 = error exit."""
    argv = (C.c_char_p * (len(args) + 1))(b"paris_b200", *[a.encode() for a in args])
    out, msg = Options(), C.create_string_buffer(4096)
    rc = _lib().paris_b200_io_parse_options(len(args) + 1, argv, C.byref(out), msg, len(msg))
    return rc, out, msg.value.decode()


def source_walk(directory: str, angle_file: str | None, quality: int, capacity: int = 4096):
    """Drain a projection source (needs a CUDA device: host projections are pinned).  Returns
    (idx, phi, angle_from_file, first_sample) arrays, one entry per projection handed out."""
    idx, phi = np.zeros(capacity, np.uint32), np.zeros(capacity, np.float32)
    from_file, first = np.zeros(capacity, np.int32), np.zeros(capacity, np.float32)
    e = _err()
    n = _lib().paris_b200_io_source_walk(directory.encode(), int(angle_file is not None), (angle_file or "").encode(),
                                         quality, idx.ctypes.data_as(_u32p), phi.ctypes.data_as(_fp),
                                         from_file.ctypes.data_as(C.POINTER(C.c_int)), first.ctypes.data_as(_fp),
                                         capacity, e, len(e))
    if n < 0:
        raise IoError(e.value.decode())
    return idx[:n].copy(), phi[:n].copy(), from_file[:n].copy(), first[:n].copy()


def scan_index(directory: str, angle_file: str | None, quality: int, capacity: int = 4096):
    """The scan as the group driver indexes it (no decoding, no device): (idx, phi, angle_from_file) per kept
    projection and (dim_x, dim_y, every file holds 16-bit samples)."""
    idx, phi = np.zeros(capacity, np.uint32), np.zeros(capacity, np.float32)
    from_file, scan3 = np.zeros(capacity, np.int32), np.zeros(3, np.uint32)
    e = _err()
    n = _lib().paris_b200_io_scan_index(directory.encode(), int(angle_file is not None), (angle_file or "").encode(),
                                        quality, idx.ctypes.data_as(_u32p), phi.ctypes.data_as(_fp),
                                        from_file.ctypes.data_as(C.POINTER(C.c_int)), capacity,
                                        scan3.ctypes.data_as(_u32p), e, len(e))
    if n < 0:
        raise IoError(e.value.decode())
    return idx[:n].copy(), phi[:n].copy(), from_file[:n].copy(), (int(scan3[0]), int(scan3[1]), bool(scan3[2]))


def scan_frame(directory: str, quality: int, i: int, dim_x: int, dim_y: int):
    """Projection i of the scan: (widened to float or None, as 16-bit counts or None)"""
    f32, u16 = np.zeros((dim_y, dim_x), np.float32), np.zeros((dim_y, dim_x), np.uint16)
    e = _err()
    got = _lib().paris_b200_io_scan_frame(directory.encode(), quality, i, f32.ctypes.data_as(_fp),
                                          u16.ctypes.data_as(C.POINTER(C.c_uint16)), e, len(e))
    if got < 0:
        raise IoError(e.value.decode())
    return (f32 if got & 1 else None), (u16 if got & 2 else None)
