"""The C++ host layer (namespace paris::b200 behind the reference's backend contract)."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle
from paris_b200 import capi, dropin

from cases import MAX_ABS_TOL, RMSE_TOL, both_det, coarse_volume, contrast, errors, shepp_logan, to_capi_vol

pytestmark = pytest.mark.gpu


def test_cpp_loop_matches_oracle(port):
    """paris_b200/cpp: load -> weight -> filter -> backproject per projection, lazily fused, one slab."""
    n, n_proj, k = 96, 48, 48
    odet, det = both_det(n, n, n_proj=n_proj)
    ovol = coarse_volume(odet, k)
    stack = shepp_logan(odet, n_proj)
    ref, _ = port.reconstruct(stack, (k, k, k), odet, ovol)
    out = np.zeros((k, k, k), np.float32)
    dropin.reconstruct(stack, n_proj, det, to_capi_vol(ovol), out, (k, k, k))
    mx, rms = errors(out, ref, contrast(n_proj))
    assert mx <= MAX_ABS_TOL and rms <= RMSE_TOL


@pytest.mark.parametrize("num_slabs", [2, 3])
def test_cpp_loop_slabs_reassemble(port, num_slabs):
    """z-slabs (remainder on the last) written at their offsets equal the one-piece reconstruction."""
    n, n_proj = 64, 24
    odet, det = both_det(n, 50, n_proj=n_proj)
    ovol = port.calculate_volume_geometry(odet)
    dims = (ovol.dim_x, ovol.dim_y, ovol.dim_z)
    stack = shepp_logan(odet, n_proj)
    whole = np.zeros((dims[2], dims[1], dims[0]), np.float32)
    dropin.reconstruct(stack, n_proj, det, to_capi_vol(ovol), whole, dims)
    parts = np.zeros_like(whole)
    for s in range(num_slabs):
        dropin.reconstruct(stack, n_proj, det, to_capi_vol(ovol), parts, dims, slab_id=s, num_slabs=num_slabs)
    assert np.array_equal(parts, whole)
    ref, _ = port.reconstruct(stack, whole.shape, odet, ovol)
    mx, rms = errors(parts, ref, contrast(n_proj))
    assert mx <= MAX_ABS_TOL and rms <= RMSE_TOL


@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(oracle.REF_PATH), "libparis_ref_b200.so")),
                    reason="oracle/_ref/libparis_ref_b200.so is only built where /root/reference exists")
def test_reference_wrappers_on_b200_backend():
    """The reference's UNMODIFIED weighting/filtering/backprojection/loader/make_volume wrappers compiled
    against namespace paris::b200 (oracle/Makefile target ref_b200) vs. the same wrappers on the
    reference's OpenMP backend."""
    n, n_proj = 64, 32
    odet, det = both_det(n, n, n_proj=n_proj)
    R = oracle.Reference()
    ovol = R.calculate_volume_geometry(odet)
    stack = shepp_logan(odet, n_proj)
    ref, _ = R.reconstruct(stack, (ovol.dim_z, ovol.dim_y, ovol.dim_x), odet, ovol)

    capi.lib()
    L = C.CDLL(os.path.join(os.path.dirname(oracle.REF_PATH), "libparis_ref_b200.so"))
    out = np.zeros_like(ref)
    vol = to_capi_vol(ovol)
    roi = capi.Roi()
    rc = L.paris_ref_b200_reconstruct(C.c_void_p(stack.ctypes.data), C.c_uint32(n_proj), C.c_void_p(out.ctypes.data),
                                      C.c_uint32(vol.dim_x), C.c_uint32(vol.dim_y), C.c_uint32(vol.dim_z),
                                      C.byref(det), C.byref(vol), C.c_int(0), C.byref(roi))
    assert rc == 0
    mx, rms = errors(out, ref, contrast(n_proj))
    assert mx <= MAX_ABS_TOL and rms <= RMSE_TOL
