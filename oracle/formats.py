"""oracle/formats.py -- file formats of PARIS restated in numpy.  TEST INFRASTRUCTURE, not product code.

* ``write_his`` / ``read_his``: the HIS frame format exactly as /root/reference/src/his.cpp:42-198 reads it
  (68-byte header read field by field :113-126, per-frame image header skipped :150-153, sample types :67-75).
  The reference has no HIS writer; ``write_his`` exists to make test inputs and is pinned by feeding its files to
  the reference's own reader (``RefIO.his_load``).
* ``read_ddbvf`` / ``ddbvf_header``: the DDBVF container as /root/reference/src/ddbvf.cpp:73-101 writes it.
* ``RefIO``: ctypes binding of the reference's unmodified I/O chain compiled into oracle/_ref/libparis_ref.so
  (oracle/ref_io_api.cpp).
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import struct
import tempfile

import numpy as np

from . import REF_PATH, build, have_ref

HIS_TYPES = {2: np.uint8, 4: np.uint16, 32: np.uint32, 64: np.float64, 128: np.float32}  # src/his.cpp:67-75


def his_header(width: int, height: int, frames: int, number_type: int, image_header_size: int = 32,
               ulx: int = 0, uly: int = 0, file_type: int = 0x7000, header_size: int = 68) -> bytes:
    """src/his.cpp:49-66 -- little endian, no padding; 68 bytes."""
    sample = np.dtype(HIS_TYPES.get(number_type, np.uint16)).itemsize
    file_size = 68 + frames * (image_header_size + width * height * sample)
    head = struct.pack("<HHHIHHHHHHHdH", file_type, header_size, 100, file_size & 0xffffffff, image_header_size,
                       ulx, uly, ulx + width - 1, uly + height - 1, frames, 0, 1000.0, number_type)
    return head + bytes(68 - len(head))


def write_his(path: str, frames: np.ndarray, number_type: int, image_header_size: int = 32, ulx: int = 0,
              uly: int = 0, **header_overrides) -> None:
    """frames: (n, height, width); samples are stored as HIS_TYPES[number_type]."""
    frames = np.asarray(frames)
    n, h, w = frames.shape
    data = frames.astype(HIS_TYPES[number_type]) if number_type in HIS_TYPES else frames.astype(np.uint16)
    with open(path, "wb") as f:
        f.write(his_header(w, h, n, number_type, image_header_size, ulx, uly, **header_overrides))
        for i in range(n):
            f.write(bytes([0xAB]) * image_header_size)   # skipped by the reader (src/his.cpp:150-153)
            f.write(np.ascontiguousarray(data[i]).tobytes())


def read_his(path: str) -> np.ndarray:
    """(n, height, width) float32, the conversion of src/his.cpp:98-99; empty array for an invalid file."""
    raw = open(path, "rb").read()
    if len(raw) < 68:
        return np.zeros((0, 0, 0), np.float32)
    (file_type, header_size, _ver, _size, img_hdr, ulx, uly, brx, bry, n, _corr, _t, number_type) = \
        struct.unpack_from("<HHHIHHHHHHHdH", raw, 0)
    if file_type != 0x7000 or header_size != 68 or number_type not in HIS_TYPES:
        return np.zeros((0, 0, 0), np.float32)
    w, h = brx - ulx + 1, bry - uly + 1
    dt = np.dtype(HIS_TYPES[number_type])
    out = np.zeros((n, h, w), np.float32)
    pos = 68
    for i in range(n):
        pos += img_hdr
        out[i] = np.frombuffer(raw, dt, w * h, pos).reshape(h, w).astype(np.float32)
        pos += w * h * dt.itemsize
    return out


def ddbvf_header(dim_x: int, dim_y: int, dim_z: int) -> bytes:
    """src/ddbvf.cpp:47-61, :73-101: id (u32), version (int, FOUR bytes), dims, offset = 8, 8 zero bytes."""
    return struct.pack("<IiIIII", 0xEFDDDAFA, 0x0010, dim_x, dim_y, dim_z, 8) + bytes(8)


def read_ddbvf(path: str) -> np.ndarray:
    """(dim_z, dim_y, dim_x) float32."""
    raw = open(path, "rb").read()
    magic, version, dx, dy, dz, off = struct.unpack_from("<IiIIII", raw, 0)
    assert magic == 0xEFDDDAFA and version == 0x0010, "not a ddbvf file"
    start = 24 + off
    return np.frombuffer(raw, np.float32, dx * dy * dz, start).reshape(dz, dy, dx).copy()


class RefIO:
    """The reference's own his::load / ddbvf::create+write / read_directory / source, from oracle/_ref.
    A private copy of the library per instance: the reference's source keeps its frame counter in a thread_local
    static (src/source.cpp:92), so one instance serves ONE source walk."""

    def __init__(self):
        build()
        if not have_ref():
            raise FileNotFoundError(REF_PATH)
        tmp = tempfile.NamedTemporaryFile(prefix="libparis_ref_io_", suffix=".so", delete=False)
        with open(REF_PATH, "rb") as src:
            shutil.copyfileobj(src, tmp)
        tmp.close()
        self.lib = C.CDLL(tmp.name)
        os.unlink(tmp.name)
        fp, u32p = C.POINTER(C.c_float), C.POINTER(C.c_uint32)
        L = self.lib
        L.paris_ref_his_load.argtypes = [C.c_char_p, fp, C.c_uint32, u32p, C.c_char_p, C.c_size_t]
        L.paris_ref_ddbvf_create_write.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.c_uint32, fp, C.c_uint32,
                                                   C.c_uint32, C.c_char_p, C.c_size_t]
        L.paris_ref_read_directory.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
        L.paris_ref_source_walk.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_uint32, u32p, fp, fp, C.c_uint32,
                                            C.c_char_p, C.c_size_t]
        L.paris_ref_make_tasks.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, u32p, u32p, C.c_uint32]

    def his_load(self, path: str, capacity_floats: int = 1 << 24) -> np.ndarray:
        out = np.zeros(capacity_floats, np.float32)
        dims = (C.c_uint32 * 2)()
        err = C.create_string_buffer(512)
        # capacity in frames is unknown before the call: decode with a generous float budget, then trim
        n = self.lib.paris_ref_his_load(path.encode(), out.ctypes.data_as(C.POINTER(C.c_float)), 0, dims, err, 512)
        if n < 0:
            raise OSError(err.value.decode())
        if n == 0:
            return np.zeros((0, 0, 0), np.float32)
        w, h = dims[0], dims[1]
        cap = min(n, capacity_floats // max(1, w * h))
        n2 = self.lib.paris_ref_his_load(path.encode(), out.ctypes.data_as(C.POINTER(C.c_float)), cap, dims, err, 512)
        assert n2 == n and cap == n, "raise capacity_floats"
        return out[:n * w * h].reshape(n, h, w).copy()

    def ddbvf_create_write(self, path_prefix: str, dims, vol: np.ndarray, first: int = 0) -> None:
        vol = np.ascontiguousarray(vol, np.float32)
        err = C.create_string_buffer(512)
        rc = self.lib.paris_ref_ddbvf_create_write(path_prefix.encode(), dims[0], dims[1], dims[2],
                                                   vol.ctypes.data_as(C.POINTER(C.c_float)), vol.shape[0], first, err, 512)
        if rc != 0:
            raise RuntimeError(err.value.decode())

    def read_directory(self, path: str) -> list[str]:
        out = C.create_string_buffer(1 << 16)
        err = C.create_string_buffer(512)
        n = self.lib.paris_ref_read_directory(path.encode(), out, 1 << 16, err, 512)
        if n < 0:
            raise RuntimeError(err.value.decode())
        return [p for p in out.value.decode().split("\n") if p]

    def source_walk(self, directory: str, angle_file: str | None, quality: int, capacity: int = 4096):
        idx = np.zeros(capacity, np.uint32)
        phi = np.zeros(capacity, np.float32)
        first = np.zeros(capacity, np.float32)
        err = C.create_string_buffer(512)
        n = self.lib.paris_ref_source_walk(directory.encode(), int(angle_file is not None),
                                           (angle_file or "").encode(), quality,
                                           idx.ctypes.data_as(C.POINTER(C.c_uint32)),
                                           phi.ctypes.data_as(C.POINTER(C.c_float)),
                                           first.ctypes.data_as(C.POINTER(C.c_float)), capacity, err, 512)
        if n < 0:
            raise RuntimeError(err.value.decode())
        return idx[:n].copy(), phi[:n].copy(), first[:n].copy()

    def make_tasks(self, num: int, dim_z: int, remainder: int):
        ids = np.zeros(max(num, 1), np.uint32)
        dz = np.zeros(max(num, 1), np.uint32)
        n = self.lib.paris_ref_make_tasks(num, dim_z, remainder, ids.ctypes.data_as(C.POINTER(C.c_uint32)),
                                          dz.ctypes.data_as(C.POINTER(C.c_uint32)), max(num, 1))
        return ids[:n].copy(), dz[:n].copy()
