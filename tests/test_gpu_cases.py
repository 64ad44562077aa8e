"""GPU parity, edge cases and size-independent properties (through the C ABI)."""
import glob
import os

import numpy as np
import pytest

import oracle
from paris_b200 import capi, phantom
from paris_b200.pipeline import Pipeline, angle_sin_cos

from cases import MAX_ABS_TOL, RMSE_TOL, both_det, coarse_volume, contrast, errors, shepp_logan, to_capi_vol

pytestmark = pytest.mark.gpu

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def _load(path):
    z = np.load(path)
    d = z["det"]
    args = (int(d[0]), int(d[1])) + tuple(float(x) for x in d[2:])
    g = z["vol_geo"]
    vol = capi.VolumeGeometry(int(g[0]), int(g[1]), int(g[2]), np.float32(g[3]), np.float32(g[4]), np.float32(g[5]))
    roi = capi.Roi(*[int(x) for x in z["roi"]]) if z["roi"].size else None
    return z, capi.DetectorGeometry(*args), vol, roi


@pytest.mark.parametrize("kernel", [2, 1])
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_against_reference_golden(ctx, path, kernel):
    """Outputs of the reference itself (tests/golden): natural volume, offset detector, coarse volume, ROI."""
    z, det, vol, roi = _load(path)
    stack = z["stack"]
    n_proj = stack.shape[0]
    ctx.set_option("bp_kernel", kernel)
    pl = Pipeline(ctx, det)
    d = pl.load(stack[0])
    pl.weight(d)
    assert np.array_equal(pl.download(d), z["weighted0"])
    pl.filter(d)
    assert np.abs(pl.download(d) - z["filtered0"]).max() <= 2e-6 * np.abs(z["filtered0"]).max()
    pl.release(d)
    shape = z["volume"].shape
    got = pl.reconstruct(stack, (shape[2], shape[1], shape[0]), vol, roi=roi, fused=True)
    pl.close()
    ctx.set_option("bp_kernel", 0)
    mx, rms = errors(got, z["volume"], contrast(n_proj))
    assert mx <= MAX_ABS_TOL and rms <= RMSE_TOL


def test_roi_and_slabs_are_bit_identical_crops(ctx):
    """Tiles are anchored in global voxel indices, so any ROI box or z-slab equals the crop of the one-piece
    reconstruction exactly -- the reference's own property, kept by the TMA kernel."""
    n, n_proj = 96, 40
    odet, det = both_det(n, 80, n_proj=n_proj)
    vol = capi.calculate_volume_geometry(det)
    stack = shepp_logan(odet, n_proj)
    ctx.set_option("bp_kernel", 2)
    pl = Pipeline(ctx, det)
    whole = pl.reconstruct(stack, (vol.dim_x, vol.dim_y, vol.dim_z), vol)
    roi = capi.Roi(7, 70, 18, 90, 5, 61)
    box = pl.reconstruct(stack, (63, 72, 56), vol, roi=roi)
    assert np.array_equal(box, whole[5:61, 18:90, 7:70])
    dz = vol.dim_z // 3
    for s in range(3):
        n_z = dz + (vol.dim_z % 3 if s == 2 else 0)
        slab = pl.reconstruct(stack, (vol.dim_x, vol.dim_y, n_z), vol, v_offset=s * dz)
        assert np.array_equal(slab, whole[s * dz:s * dz + n_z])
    pl.close()
    ctx.set_option("bp_kernel", 0)


@pytest.mark.parametrize("v_offset,dz", [(0, 200), (5, 150), (70, 64)])
def test_backproject_with_overlapped_download_equals_separate_calls(ctx, v_offset, dz):
    """paris_b200_backproject_stack_d2h cuts the slab into z-chunks at the kernel's tile anchors and downloads each
    chunk behind the next one's backprojection; the host volume must equal backproject_stack + vol_d2h bit for bit,
    whatever the slab's offset (chunks then start off the anchors) and for slabs of several chunks or one."""
    n, n_proj = 128, 70
    _, det = both_det(n, 300, l_px=0.4, n_proj=n_proj)
    vol = capi.calculate_volume_geometry(det)
    assert vol.dim_z >= v_offset + dz
    raw = ctx.dev_alloc(n_proj * n * 300 * 4)
    ell = phantom.scaled_ellipsoids(phantom.SHEPP_LOGAN_3D, 0.9 * phantom.fov_radius(n, 0.4, 0, 500, 500))
    ctx.phantom_project(ell, det, 0, n_proj, raw)
    stack = ctx.stack_alloc(n, 300, n_proj)
    filt = ctx.filter_create(capi.filter_size(n), 0.4)
    layout = capi.choose_stack_layout(det, vol)
    ctx.filter_to_stack_batch(raw, n * 300, n_proj, det, filt, stack, 0, layout)
    sc = np.array([angle_sin_cos(i, det) for i in range(n_proj)], dtype=np.float32)
    dims = (vol.dim_x, vol.dim_y, dz)
    ctx.set_option("bp_batch", 32)          # several launches per chunk as well
    v = ctx.volume_alloc(*dims)
    ctx.backproject_stack(stack, 0, n_proj, sc[:, 0], sc[:, 1], v, dims, v_offset, det, vol, layout=layout)
    want = np.empty((dz, vol.dim_y, vol.dim_x), np.float32)
    ctx.vol_d2h(v, want, want.size)
    ctx.volume_clear(v, *dims)
    got = capi.PinnedArray(want.shape)
    ctx.backproject_stack_d2h(stack, 0, n_proj, sc[:, 0], sc[:, 1], v, dims, v_offset, det, vol, got.ptr, layout=layout)
    same = np.array_equal(got.array, want)
    nonzero = float(np.abs(want).max())
    got.free()
    ctx.set_option("bp_batch", 256)
    ctx.volume_free(v)
    ctx.filter_destroy(filt)
    ctx.stack_free(stack)
    ctx.dev_free(raw)
    assert nonzero > 0 and same


@pytest.mark.parametrize("n_row,n_col,n_proj", [(100, 37, 9), (33, 65, 3), (250, 20, 65), (64, 64, 1)])
def test_ragged_shapes_and_batch_boundaries(ctx, port, n_row, n_col, n_proj):
    """odd detector sizes (pitch padding, partial row groups), volumes that are no multiple of the tile,
    projection counts that are no multiple of the batch (65 = 64 + 1), a single projection."""
    odet, det = both_det(n_row, n_col, n_proj=max(n_proj, 8))
    ovol = port.calculate_volume_geometry(odet)
    rng = np.random.default_rng(n_row)
    stack = (shepp_logan(odet, n_proj) + rng.normal(0, 1e-3, (n_proj, n_col, n_row))).astype(np.float32)
    shape = (ovol.dim_z, ovol.dim_y, ovol.dim_x)
    ref, _ = port.reconstruct(stack, shape, odet, ovol)
    pl = Pipeline(ctx, det)
    got = pl.reconstruct(stack, (ovol.dim_x, ovol.dim_y, ovol.dim_z), to_capi_vol(ovol))
    pl.close()
    mx, rms = errors(got, ref, contrast(max(n_proj, 8)))
    assert mx <= MAX_ABS_TOL and rms <= RMSE_TOL


def test_linearity(ctx):
    """reconstruct(a x + b y) == a reconstruct(x) + b reconstruct(y) up to float rounding."""
    n, n_proj, k = 64, 16, 32
    odet, det = both_det(n, n, n_proj=n_proj)
    vol = to_capi_vol(coarse_volume(odet, k))
    rng = np.random.default_rng(5)
    x = rng.standard_normal((n_proj, n, n)).astype(np.float32)
    y = rng.standard_normal((n_proj, n, n)).astype(np.float32)
    pl = Pipeline(ctx, det)
    rx, ry = pl.reconstruct(x, (k, k, k), vol), pl.reconstruct(y, (k, k, k), vol)
    rz = pl.reconstruct((2.0 * x - 0.5 * y).astype(np.float32), (k, k, k), vol)
    pl.close()
    scale = np.abs(rz).max()
    assert np.abs(rz - (2.0 * rx - 0.5 * ry)).max() <= 2e-5 * scale


def test_changing_the_target_flushes(ctx, port):
    """Two volumes fed alternately: every switch must flush the pending batch of the other."""
    n, n_proj = 48, 6
    odet, det = both_det(n, n, n_proj=n_proj)
    ovol = port.calculate_volume_geometry(odet)
    vol = to_capi_vol(ovol)
    stack = shepp_logan(odet, n_proj)
    pl = Pipeline(ctx, det)
    va = pl.make_volume(vol.dim_x, vol.dim_y, vol.dim_z)
    vb = pl.make_volume(vol.dim_x, vol.dim_y, vol.dim_z)
    for i in range(n_proj):
        d = pl.load(stack[i], idx=i)
        pl.backproject(d, va if i % 2 == 0 else vb, 0, vol, fused_raw=True)
        pl.release(d)
    a, b = pl.save(va), pl.save(vb)
    pl.free_volume(va)
    pl.free_volume(vb)
    pl.close()
    shape = (ovol.dim_z, ovol.dim_y, ovol.dim_x)
    ra = np.zeros(shape, np.float32)
    rb = np.zeros(shape, np.float32)
    for i in range(n_proj):
        f = port.filter(port.weight(stack[i], odet), odet)
        port.backproject(f, i, ra if i % 2 == 0 else rb, odet, ovol)
    c = contrast(n_proj)
    assert errors(a, ra, c)[0] <= MAX_ABS_TOL and errors(b, rb, c)[0] <= MAX_ABS_TOL


def test_phantom_kernel_matches_numpy(ctx):
    odet, det = both_det(72, 40, l_px=0.3, delta_s=1.5, delta_t=-0.5, n_proj=12)
    ell = phantom.scaled_ellipsoids(phantom.SHEPP_LOGAN_3D, 0.9 * phantom.fov_radius(72, 0.3, 1.5, 500, 500))
    import oracle.phantom
    want = oracle.phantom.shepp_logan_stack(72, 40, 0.3, 0.3, 1.5, -0.5, 500, 500, 12)
    d = ctx.dev_alloc(want.nbytes)
    ctx.phantom_project(ell, det, 0, 12, d)
    got = np.empty_like(want)
    for i in range(12):
        ctx.proj_d2h(d + i * 72 * 40 * 4, got[i], 72, 40)
    ctx.dev_free(d)
    assert np.abs(got - want).max() <= 1e-4 * want.max()


@pytest.mark.parametrize("n_row,n_col", [(96, 96), (100, 37), (256, 64), (1024, 16), (2048, 8), (3000, 5)])
@pytest.mark.parametrize("layout", [capi.LAYOUT_PLAIN, capi.LAYOUT_SPLIT2])
def test_filter_to_stack_layouts(ctx, port, n_row, n_col, layout):
    """The fused weight+filter kernel writing a TRANSPOSED stack slot, plain and parity-split line layout."""
    odet, det = both_det(n_row, n_col, l_px=0.2, delta_s=1.0, delta_t=-1.0)
    rng = np.random.default_rng(11)
    p = rng.standard_normal((n_col, n_row)).astype(np.float32)
    ref = port.filter(port.weight(p, odet), odet)
    f = ctx.filter_create(capi.filter_size(n_row), 0.2)
    slot_bytes, pitch = capi.stack_slot_bytes(n_row, n_col)
    d = ctx.dev_alloc(p.nbytes)
    ctx.proj_h2d(p, d, n_row, n_col)
    st = ctx.stack_alloc(n_row, n_col, 2)
    ctx.filter_to_stack(d, det, f, st, 1, layout)
    out = np.empty((n_row, pitch), np.float32)
    ctx.proj_d2h(st + slot_bytes, out, pitch, n_row)
    ctx.dev_free(d)
    ctx.stack_free(st)
    ctx.filter_destroy(f)
    if layout == capi.LAYOUT_SPLIT2:
        plain = np.empty_like(out)
        plain[:, 0::2] = out[:, :pitch // 2]
        plain[:, 1::2] = out[:, pitch // 2:]
        out = plain
    assert np.abs(out[:, :n_col].T - ref).max() <= 2e-6 * np.abs(ref).max()
    assert not out[:, n_col:].any()   # the slot's padding columns stay zero
