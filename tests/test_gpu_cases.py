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
    want = phantom.shepp_logan_stack(72, 40, 0.3, 0.3, 1.5, -0.5, 500, 500, 12)
    d = ctx.dev_alloc(want.nbytes)
    ctx.phantom_project(ell, det, 0, 12, d)
    got = np.empty_like(want)
    for i in range(12):
        ctx.proj_d2h(d + i * 72 * 40 * 4, got[i], 72, 40)
    ctx.dev_free(d)
    assert np.abs(got - want).max() <= 1e-4 * want.max()


def test_full_size_config2_fast_kernel_against_exact_kernel(ctx):
    """BASELINE config 2 at full size (512^3 from 720 x 1024^2): the TMA kernel against the exact kernel
    (which the small cases pin bit for bit to the reference), both on the same filtered stack, plus the
    size-independent sanity property that the phantom's plateau sits at density * N / (8 pi)."""
    n, n_proj, k = 1024, 720, 512
    l_px = 0.2
    det = capi.DetectorGeometry(n, n, l_px, l_px, 0, 0, 500, 500, 360.0 / n_proj)
    nat = capi.calculate_volume_geometry(det)
    f32 = np.float32
    vol = capi.VolumeGeometry(k, k, k, f32(nat.l_vx_x * nat.dim_x / k), f32(nat.l_vx_y * nat.dim_y / k),
                              f32(nat.l_vx_z * nat.dim_z / k))
    raw = ctx.dev_alloc(n_proj * n * n * 4)
    ell = phantom.scaled_ellipsoids(phantom.SHEPP_LOGAN_3D, 0.9 * phantom.fov_radius(n, l_px, 0, 500, 500))
    ctx.phantom_project(ell, det, 0, n_proj, raw)
    slot_bytes, _ = capi.stack_slot_bytes(n, n)
    stack = ctx.stack_alloc(n, n, n_proj)
    filt = ctx.filter_create(capi.filter_size(n), l_px)
    layout = capi.choose_stack_layout(det, vol)
    assert layout == capi.LAYOUT_SPLIT2
    ctx.filter_to_stack_batch(raw, n * n, n_proj, det, filt, stack, 0, layout)
    sc = np.array([angle_sin_cos(i, det) for i in range(n_proj)], dtype=np.float32)
    out = {}
    for kernel in (2, 1):
        ctx.set_option("bp_kernel", kernel)
        v = ctx.volume_alloc(k, k, k)
        ctx.backproject_stack(stack, 0, n_proj, sc[:, 0], sc[:, 1], v, (k, k, k), 0, det, vol, layout=layout)
        out[kernel] = np.empty((k, k, k), np.float32)
        ctx.vol_d2h(v, out[kernel], k ** 3)
        ctx.volume_free(v)
    ctx.set_option("bp_kernel", 0)
    ctx.filter_destroy(filt)
    ctx.stack_free(stack)
    ctx.dev_free(raw)
    c = contrast(n_proj)
    mx, rms = errors(out[2], out[1], c)
    print(f"config 2 full size: fast vs exact kernel max {mx:.2e} C, rmse {rms:.2e} C")
    assert mx <= MAX_ABS_TOL and rms <= RMSE_TOL
    # brain tissue of the phantom (density 1.0 - 0.8 = 0.2) a few voxels off the centre, away from the ventricles
    plateau = out[2][k // 2, k // 2 - 40:k // 2 - 30, k // 2 - 4:k // 2 + 4].mean()
    assert abs(plateau / (0.2 * c) - 1.0) < 0.05


def test_full_size_config4_roi_offset_detector(ctx):
    """BASELINE config 4 at full size: 1024^3 region of interest of the 2248 x 2248 x 2060 natural volume from 2880
    projections of a 2048^2 detector shifted by 100 pixels.  Properties checked (no CPU oracle finishes this size):
    the production kernel agrees with the exact kernel -- pinned bit for bit to the reference on the small cases --
    on bands of the ROI at its bottom, middle and top, and the ROI reconstruction is a bit-identical crop whether it
    is computed in one piece or band by band (what z-slab tasks and the chunked download rely on)."""
    n, n_proj, k = 2048, 2880, 1024
    l_px = 0.1
    det = capi.DetectorGeometry(n, n, l_px, l_px, 100.0, 0, 500, 500, 360.0 / n_proj)
    nat = capi.calculate_volume_geometry(det)
    assert (nat.dim_x, nat.dim_y, nat.dim_z) == (2248, 2248, 2060)
    roi = capi.Roi(612, 1636, 612, 1636, 518, 1542)
    reg = capi.apply_roi(nat, roi)
    assert (reg.dim_x, reg.dim_y, reg.dim_z) == (k, k, k)
    layout = capi.choose_stack_layout(det, nat)
    assert layout == capi.LAYOUT_PLAIN
    stack = ctx.stack_alloc(n, n, n_proj)                               # 48 GB
    filt = ctx.filter_create(capi.filter_size(n), l_px)
    ell = phantom.scaled_ellipsoids(phantom.SHEPP_LOGAN_3D, 0.9 * phantom.fov_radius(n, l_px, 100.0, 500, 500))
    raw = ctx.dev_alloc(64 * n * n * 4)
    for first in range(0, n_proj, 64):                                  # raw projections never exist all at once
        ctx.phantom_project(ell, det, first, 64, raw)
        ctx.filter_to_stack_batch(raw, n * n, 64, det, filt, stack, first, layout)
    ctx.dev_free(raw)
    sc = np.array([angle_sin_cos(i, det) for i in range(n_proj)], dtype=np.float32)

    ctx.set_option("bp_kernel", 2)
    v = ctx.volume_alloc(k, k, k)
    e0 = ctx.event()
    ctx.backproject_stack(stack, 0, n_proj, sc[:, 0], sc[:, 1], v, (k, k, k), 0, det, nat, roi=roi, layout=layout)
    e1 = ctx.event()
    ms = ctx.elapsed_ms(e0, e1)
    print(f"config 4 full size: {k ** 3 * n_proj / ms / 1e6:.0f} GUPS ({ms:.0f} ms)")
    full = np.empty((k, k, k), np.float32)
    ctx.vol_d2h(v, full, k ** 3)
    ctx.volume_free(v)

    c = contrast(n_proj)
    band = 4
    for z in (0, 509, k - band):                                        # bands of `band` slices at these ROI offsets
        got = {}
        for kernel in (2, 1):
            ctx.set_option("bp_kernel", kernel)
            vb = ctx.volume_alloc(k, k, band)
            ctx.backproject_stack(stack, 0, n_proj, sc[:, 0], sc[:, 1], vb, (k, k, band), z, det, nat, roi=roi,
                                  layout=layout)
            got[kernel] = np.empty((band, k, k), np.float32)
            ctx.vol_d2h(vb, got[kernel], band * k * k)
            ctx.volume_free(vb)
        assert np.array_equal(got[2], full[z:z + band]), f"band at {z} is not a crop of the one-piece ROI"
        mx, rms = errors(got[2], got[1], c)
        print(f"config 4 band at z={z}: fast vs exact kernel max {mx:.2e} C, rmse {rms:.2e} C")
        assert mx <= MAX_ABS_TOL and rms <= RMSE_TOL
    ctx.set_option("bp_kernel", 0)
    ctx.filter_destroy(filt)
    ctx.stack_free(stack)
    assert np.isfinite(full).all() and full.max() > 0.5 * 1.0 * c * 0.2


def test_full_size_config5_slabs_stream_over_one_filtered_stack(ctx):
    """BASELINE config 5 geometry at full size: the 2048^3 region (z in [5, 2053) of the 2048 x 2048 x 2058 natural
    volume) from 2880 projections of a 2048^2 detector, reconstructed slab by slab from ONE filtered stack (the
    reference re-reads and re-filters the whole scan for every sub-volume, src/main.cpp:93-105).  Two adjacent
    128-slice slabs of the 16 are computed here: streamed separately they must be bit-identical to the same 256
    slices computed in one piece, and bands of them must agree with the exact kernel."""
    n, n_proj, k = 2048, 2880, 2048
    l_px = 0.1
    det = capi.DetectorGeometry(n, n, l_px, l_px, 0, 0, 500, 500, 360.0 / n_proj)
    nat = capi.calculate_volume_geometry(det)
    assert (nat.dim_x, nat.dim_y, nat.dim_z) == (2048, 2048, 2058)
    roi = capi.Roi(0, 2047, 0, 2047, 5, 2053)
    reg = capi.apply_roi(nat, roi)
    assert (reg.dim_x, reg.dim_y, reg.dim_z) == (k, k, k)
    layout = capi.choose_stack_layout(det, nat)
    stack = ctx.stack_alloc(n, n, n_proj)                               # 48 GB, filled once
    filt = ctx.filter_create(capi.filter_size(n), l_px)
    ell = phantom.scaled_ellipsoids(phantom.SHEPP_LOGAN_3D, 0.9 * phantom.fov_radius(n, l_px, 0, 500, 500))
    raw = ctx.dev_alloc(64 * n * n * 4)
    for first in range(0, n_proj, 64):
        ctx.phantom_project(ell, det, first, 64, raw)
        ctx.filter_to_stack_batch(raw, n * n, 64, det, filt, stack, first, layout)
    ctx.dev_free(raw)
    sc = np.array([angle_sin_cos(i, det) for i in range(n_proj)], dtype=np.float32)
    slab = 128                                                          # 16 slabs of 128 slices

    def reconstruct(z_first, dz, kernel):
        ctx.set_option("bp_kernel", kernel)
        v = ctx.volume_alloc(k, k, dz)
        e0 = ctx.event()
        ctx.backproject_stack(stack, 0, n_proj, sc[:, 0], sc[:, 1], v, (k, k, dz), z_first, det, nat, roi=roi, layout=layout)
        e1 = ctx.event()
        ms = ctx.elapsed_ms(e0, e1)
        out = np.empty((dz, k, k), np.float32)
        ctx.vol_d2h(v, out, dz * k * k)
        ctx.volume_free(v)
        return out, ms

    a, ms_a = reconstruct(7 * slab, slab, 2)
    b, _ = reconstruct(8 * slab, slab, 2)
    print(f"config 5 slab of {slab} slices: {k * k * slab * n_proj / ms_a / 1e6:.0f} GUPS ({ms_a:.0f} ms)")
    both, _ = reconstruct(7 * slab, 2 * slab, 2)
    assert np.array_equal(both[:slab], a) and np.array_equal(both[slab:], b)
    c = contrast(n_proj)
    for z, src in ((7 * slab, a), (9 * slab - 2, b)):
        exact, _ = reconstruct(z, 2, 1)
        got = src[z - (7 * slab if src is a else 8 * slab):][:2]
        mx, rms = errors(got, exact, c)
        print(f"config 5 band at z={z}: fast vs exact kernel max {mx:.2e} C, rmse {rms:.2e} C")
        assert mx <= MAX_ABS_TOL and rms <= RMSE_TOL
    ctx.set_option("bp_kernel", 0)
    ctx.filter_destroy(filt)
    ctx.stack_free(stack)


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
