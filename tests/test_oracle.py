"""The CPU oracle, pinned: the plain-C restatement (oracle/fdk_oracle.c) against
 (a) golden outputs of the reference itself (tests/golden/*.npz, made by tests/golden/make_golden.py from
     oracle/_ref/libparis_ref.so = the unmodified /root/reference sources),
 (b) the reference build live, where it exists,
 (c) known answers that follow from the reference's own formulas (it ships no tests or vectors)."""
import glob
import os

import numpy as np
import pytest

import oracle
import oracle.phantom as phantom

from cases import both_det, coarse_volume, shepp_logan

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def _load(path):
    z = np.load(path)
    d = z["det"]
    det = oracle.DetectorGeometry(int(d[0]), int(d[1]), *[float(x) for x in d[2:]])
    g = z["vol_geo"]
    vol = oracle.VolumeGeometry(int(g[0]), int(g[1]), int(g[2]), np.float32(g[3]), np.float32(g[4]), np.float32(g[5]))
    roi = oracle.Roi(*[int(x) for x in z["roi"]]) if z["roi"].size else None
    return z, det, vol, roi


def test_golden_present():
    assert len(GOLDEN) >= 4


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_port_reproduces_reference_golden(port, path):
    z, det, vol, roi = _load(path)
    stack = z["stack"]
    assert np.array_equal(port.weight(stack[0], det), z["weighted0"])
    assert np.array_equal(port.filter(z["weighted0"], det), z["filtered0"])
    assert np.array_equal(port.make_filter(port.filter_size(det.n_row), det.l_px_row), z["k"])
    got, _ = port.reconstruct(stack, z["volume"].shape, det, vol, roi=roi)
    assert np.array_equal(got, z["volume"])


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref is only built where /root/reference exists")
@pytest.mark.parametrize("n_row,n_col,delta_s,delta_t", [(32, 24, 0.0, 0.0), (50, 37, 1.5, -2.0), (96, 64, -3.0, 0.5)])
def test_port_equals_reference_live(port, n_row, n_col, delta_s, delta_t):
    odet, _ = both_det(n_row, n_col, delta_s=delta_s, delta_t=delta_t, n_proj=12)
    ref = oracle.Reference()
    vg, vr = port.calculate_volume_geometry(odet), ref.calculate_volume_geometry(odet)
    assert [getattr(vg, f[0]) for f in vg._fields_] == [getattr(vr, f[0]) for f in vr._fields_]
    rng = np.random.default_rng(7)
    stack = rng.standard_normal((12, n_col, n_row)).astype(np.float32)
    assert np.array_equal(port.weight(stack[0], odet), ref.weight(stack[0], odet))
    assert np.array_equal(port.filter(stack[0], odet), ref.filter(stack[0], odet))
    shape = (vg.dim_z, vg.dim_y, vg.dim_x)
    a, _ = port.reconstruct(stack, shape, odet, vg)
    b, _ = ref.reconstruct(stack, shape, odet, vr)
    assert np.array_equal(a, b)


def test_fft_standin_matches_numpy(port):
    """oracle/fft_shim.c behind the FFTW symbols: filter == irfft(rfft(padded) * K) / ... in float64."""
    odet, _ = both_det(100, 7, l_px=0.25)
    rng = np.random.default_rng(3)
    p = rng.standard_normal((7, 100)).astype(np.float32)
    n = port.filter_size(100)
    assert n == 256
    k = port.make_filter(n, 0.25).astype(np.float64)
    padded = np.zeros((7, n))
    padded[:, :100] = p
    want = np.fft.irfft(np.fft.rfft(padded, axis=1) * k, n=n, axis=1)[:, :100]
    got = port.filter(p, odet)
    assert np.abs(got - want).max() <= 2e-6 * np.abs(want).max()


@pytest.mark.parametrize("size,tau", [(64, 0.4), (512, 0.2), (4096, 0.1)])
def test_filter_table_known_answers(port, size, tau):
    """K = tau*|DFT(r)| with r(0) = 1/(8 tau^2), r(odd j) = -1/(2 j^2 pi^2 tau^2), r(even) = 0
    (src/openmp/filtering.cpp:63-70) is the ramp |k| / (2 N tau) up to the truncation of the taps."""
    k = port.make_filter(size, tau).astype(np.float64)
    assert k.shape == (size // 2 + 1,)
    j = np.arange(-(size - 2) // 2, -(size - 2) // 2 + size)
    jj = np.where(j == 0, 1, j).astype(np.float64)
    r = np.where(j == 0, 1.0 / (8 * tau * tau), np.where(j % 2 == 0, 0.0, -1.0 / (2.0 * jj ** 2 * np.pi ** 2 * tau ** 2)))
    want = tau * np.abs(np.fft.rfft(r))
    np.testing.assert_allclose(k, want, rtol=2e-5, atol=1e-7 * want.max())
    ramp = np.arange(size // 2 + 1) / (2.0 * size * tau)
    assert np.abs(k - ramp).max() <= 2.0 / (size * size * tau) * 4 + 1e-3 * ramp.max()


def test_weight_known_answers(port):
    """w(s,t) = d_sd / sqrt(d_sd^2 + h_s^2 + v_t^2): 1 on the central ray, symmetric for a centred detector."""
    odet, _ = both_det(65, 33, l_px=0.5)
    w = port.weight(np.ones((33, 65), np.float32), odet)
    assert w.max() <= 1.0
    assert w[16, 32] == w.max()
    assert abs(w[16, 32] - 1.0) < 1e-6
    np.testing.assert_allclose(w, w[::-1, ::-1], rtol=0, atol=1e-6)
    h, v = (0 - 32) * 0.5, (0 - 16) * 0.5
    assert abs(w[0, 0] - 1000.0 / np.sqrt(1000.0 ** 2 + h * h + v * v)) < 1e-6


@pytest.mark.parametrize("n,l_px,delta_s,expect", [
    (256, 0.4, 0.0, (256, 256, 256)),
    (1024, 0.2, 0.0, (1024, 1024, 1029)),
    (2048, 0.1, 0.0, (2048, 2048, 2058)),
    (2048, 0.1, 100.0, (2248, 2248, 2060)),
])
def test_volume_geometry_table(port, n, l_px, delta_s, expect):
    """calculate_volume_geometry derives the volume from the detector."""
    odet, _ = both_det(n, n, l_px=l_px, delta_s=delta_s)
    v = port.calculate_volume_geometry(odet)
    assert (v.dim_x, v.dim_y, v.dim_z) == expect
    assert v.l_vx_x == v.l_vx_y == v.l_vx_z


def test_apply_roi_semantics(port):
    """dim = x2 - x1, +1 iff x1 == 0; an invalid or oversized ROI leaves the geometry alone (src/geometry.cpp:86-130)."""
    v = oracle.VolumeGeometry(100, 100, 80, 0.5, 0.5, 0.5)
    g = port.apply_roi(v, oracle.Roi(10, 40, 0, 50, 5, 25))
    assert (g.dim_x, g.dim_y, g.dim_z) == (30, 51, 20)
    g = port.apply_roi(v, oracle.Roi(40, 10, 0, 50, 5, 25))
    assert (g.dim_x, g.dim_y, g.dim_z) == (100, 100, 80)
    g = port.apply_roi(v, oracle.Roi(0, 100, 0, 50, 5, 25))   # 101 > 100
    assert (g.dim_x, g.dim_y, g.dim_z) == (100, 100, 80)


def test_unit_sphere_plateau(port):
    """The reference's output is not normalised by the angular step: a unit-density sphere reconstructs to
    N_proj / (8 pi)."""
    n, n_proj = 64, 64
    odet, _ = both_det(n, n, n_proj=n_proj)
    vg = port.calculate_volume_geometry(odet)
    r = phantom.fov_radius(n, 0.4, 0, 500, 500)
    ang = np.float32(360.0 / n_proj) * np.arange(n_proj, dtype=np.float32)
    st = phantom.project(phantom.scaled_ellipsoids(phantom.UNIT_SPHERE, 0.5 * r), n, n, 0.4, 0.4, 0, 0, 500, 500, ang)
    vol, _ = port.reconstruct(st, (vg.dim_z, vg.dim_y, vg.dim_x), odet, vg)
    c = vol[vg.dim_z // 2, vg.dim_y // 2 - 3:vg.dim_y // 2 + 3, vg.dim_x // 2 - 3:vg.dim_x // 2 + 3].mean()
    assert abs(c / (n_proj / (8 * np.pi)) - 1.0) < 5e-3


def test_slabs_and_roi_blocks_are_bit_identical_crops(port):
    """z-slabs (v_offset) and ROI boxes reproduce the corresponding crop of the one-piece reconstruction
    exactly -- the property the multi-GPU split relies on."""
    n, n_proj = 40, 10
    odet, _ = both_det(n, 36, n_proj=n_proj)
    vg = port.calculate_volume_geometry(odet)
    stack = shepp_logan(odet, n_proj)
    filtered = [port.filter(port.weight(s, odet), odet) for s in stack]
    whole = np.zeros((vg.dim_z, vg.dim_y, vg.dim_x), np.float32)
    for i, f in enumerate(filtered):
        port.backproject(f, i, whole, odet, vg)
    info = port.make_subvolume_information(vg, 3)
    assert info.dim_z == vg.dim_z // 3 and info.remainder == vg.dim_z % 3
    for s in range(3):
        dz = info.dim_z + (info.remainder if s == 2 else 0)
        slab = np.zeros((dz, vg.dim_y, vg.dim_x), np.float32)
        for i, f in enumerate(filtered):
            port.backproject(f, i, slab, odet, vg, v_offset=s * info.dim_z)
        assert np.array_equal(slab, whole[s * info.dim_z:s * info.dim_z + dz])
    roi = oracle.Roi(5, 25, 8, 30, 4, 20)
    g = port.apply_roi(vg, roi)
    box = np.zeros((g.dim_z, g.dim_y, g.dim_x), np.float32)
    for i, f in enumerate(filtered):
        port.backproject(f, i, box, odet, vg, roi=roi)
    assert np.array_equal(box, whole[4:20, 8:30, 5:25])


def test_coarse_volume_helper_keeps_field_of_view(port):
    odet, _ = both_det(128, 128)
    nat = port.calculate_volume_geometry(odet)
    c = coarse_volume(odet, 64)
    assert abs(c.l_vx_x * 64 - nat.l_vx_x * nat.dim_x) < 1e-4
