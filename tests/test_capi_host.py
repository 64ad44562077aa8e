"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU, exports exactly what
include/paris_b200.h declares, refuses to compute without an sm_100a device (no fallback), and its
host-side geometry arithmetic equals the reference's."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle
from paris_b200 import capi, dropin, pipeline

from cases import both_det

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "paris_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(paris_b200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = capi.lib()
    names = _header_symbols()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/paris_b200.h but not exported"


def test_binding_covers_the_header():
    assert sorted(capi.SIGNATURES) == _header_symbols()


def test_dropin_library_loads():
    L = dropin.lib()
    for n in ("paris_b200_dropin_reconstruct", "paris_b200_dropin_context", "paris_b200_dropin_set_device"):
        assert hasattr(L, n)


def test_no_silent_fallback_without_a_gpu():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    with pytest.raises(capi.Error) as e:
        capi.Context(0)
    assert e.value.code in (capi.ECUDA, capi.EINVAL)
    out = np.zeros((4, 4, 4), np.float32)
    det = capi.DetectorGeometry(8, 8, 1, 1, 0, 0, 100, 100, 1)
    vol = capi.VolumeGeometry(4, 4, 4, 1, 1, 1)
    with pytest.raises(capi.Error):
        dropin.reconstruct(np.zeros((1, 8, 8), np.float32), 1, det, vol, out, (4, 4, 4))
    assert not out.any()


@pytest.mark.parametrize("n_row,n_col,l_px,delta_s,delta_t,d_so,d_od", [
    (256, 256, 0.4, 0.0, 0.0, 500.0, 500.0), (1024, 1024, 0.2, 0.0, 0.0, 500.0, 500.0),
    (2048, 2048, 0.1, 100.0, 0.0, 500.0, 500.0), (100, 37, 0.3, -3.5, 2.0, 300.0, 700.0),
    (640, 480, 0.127, 0.0, 12.0, 1000.0, 250.0)])
def test_geometry_arithmetic_equals_reference(port, n_row, n_col, l_px, delta_s, delta_t, d_so, d_od):
    odet, det = both_det(n_row, n_col, l_px=l_px, delta_s=delta_s, delta_t=delta_t, d_so=d_so, d_od=d_od)
    a, b = capi.calculate_volume_geometry(det), port.calculate_volume_geometry(odet)
    assert [getattr(a, f[0]) for f in a._fields_] == [getattr(b, f[0]) for f in b._fields_]
    assert capi.filter_size(n_row) == port.filter_size(n_row)
    for roi in [(0, 10, 0, 10, 0, 10), (5, 50, 6, 30, 7, 20), (50, 5, 6, 30, 7, 20), (0, b.dim_x, 0, 5, 0, 5)]:
        ra = capi.apply_roi(a, capi.Roi(*roi))
        rb = port.apply_roi(b, oracle.Roi(*roi))
        assert (ra.dim_x, ra.dim_y, ra.dim_z) == (rb.dim_x, rb.dim_y, rb.dim_z)


def test_filter_size_is_twice_next_power_of_two():
    for n, want in [(1, 2), (2, 4), (3, 8), (64, 128), (100, 256), (1024, 2048), (1025, 4096), (4096, 8192)]:
        assert capi.filter_size(n) == want


def test_forced_slab_split_needs_no_device(port):
    """make_subvolume_information with an explicit slab count is pure arithmetic (ctx may be NULL)."""
    v = capi.VolumeGeometry(64, 64, 50, 1, 1, 1)
    det = capi.DetectorGeometry(64, 64, 1, 1, 0, 0, 100, 100, 1)
    out = capi.SubvolumeInfo()
    capi.check(capi.lib().paris_b200_make_subvolume_information(None, C.byref(v), C.byref(det), 8, C.byref(out)))
    assert (out.dim_x, out.dim_y, out.dim_z, out.remainder, out.num) == (64, 64, 6, 2, 8)
    ref = port.make_subvolume_information(oracle.VolumeGeometry(64, 64, 50, 1, 1, 1), 8)
    assert (ref.dim_z, ref.remainder, ref.num) == (out.dim_z, out.remainder, out.num)


def test_stack_slot_layout():
    nbytes, pitch = capi.stack_slot_bytes(1024, 1000)
    assert pitch == 1024 and nbytes == 1024 * 1024 * 4
    nbytes, pitch = capi.stack_slot_bytes(300, 33)
    assert pitch == 64 and nbytes == 300 * 64 * 4


def test_wrapper_constants_match_reference_arithmetic(port):
    """pipeline.weight_constants / angle_sin_cos restate src/weighting.cpp:37-42 and src/backprojection.cpp:53-63
    in float32: feeding them to the oracle's backend-level formula must reproduce the oracle's wrapper."""
    odet, det = both_det(50, 30, l_px=0.3, delta_s=2.0, delta_t=-1.0)
    h_min, v_min, d_sd = pipeline.weight_constants(det)
    f32 = np.float32
    s, t = np.meshgrid(np.arange(50, dtype=f32), np.arange(30, dtype=f32))
    l = f32(0.3)
    h_s = (l / f32(2) + s * l) + f32(h_min)
    v_t = (l / f32(2) + t * l) + f32(v_min)
    w = f32(d_sd) / np.sqrt(f32(d_sd) * f32(d_sd) + h_s * h_s + v_t * v_t, dtype=f32)
    got = port.weight(np.ones((30, 50), f32), odet)
    assert np.abs(got - w).max() <= 1.2e-7
    sn, cs = pipeline.angle_sin_cos(7, det)
    phi = np.float32(7) * np.float32(det.delta_phi) * (np.float32(np.pi) / np.float32(180))
    assert abs(sn - np.sin(np.float64(phi))) < 1e-7 and abs(cs - np.cos(np.float64(phi))) < 1e-7
