"""GPU parity: the sm_100a kernels, called through the C ABI, against the CPU oracle on the same inputs."""
import numpy as np
import pytest

import oracle
from paris_b200 import capi
from paris_b200.pipeline import Pipeline, weight_constants

from cases import MAX_ABS_TOL, RMSE_TOL, both_det, coarse_volume, contrast, errors, shepp_logan, to_capi_vol

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_row,n_col,delta_s,delta_t", [(64, 48, 0.0, 0.0), (100, 37, 3.0, -2.0), (256, 256, 0.0, 0.0)])
def test_weight_bit_exact(ctx, port, n_row, n_col, delta_s, delta_t):
    odet, det = both_det(n_row, n_col, delta_s=delta_s, delta_t=delta_t)
    rng = np.random.default_rng(1)
    p = rng.standard_normal((n_col, n_row)).astype(np.float32)
    pl = Pipeline(ctx, det)
    d = pl.load(p)
    pl.weight(d)
    got = pl.download(d)
    pl.release(d)
    assert np.array_equal(got, port.weight(p, odet))


@pytest.mark.parametrize("size,tau", [(32, 0.4), (128, 0.4), (512, 0.4), (2048, 0.2), (4096, 0.1), (8192, 0.05)])
def test_filter_table(ctx, port, size, tau):
    f = ctx.filter_create(size, tau)
    k = ctx.filter_read(f, size)
    ctx.filter_destroy(f)
    ref = port.make_filter(size, tau)
    # same float taps, transform in double on both sides, rounded to float: at most an ulp apart
    np.testing.assert_allclose(k, ref, rtol=3e-7, atol=0)


@pytest.mark.parametrize("n_row,n_col", [(16, 8), (64, 48), (100, 37), (256, 256), (1024, 33), (2048, 16), (3000, 9)])
def test_apply_filter(ctx, port, n_row, n_col):
    odet, det = both_det(n_row, n_col, l_px=0.2)
    rng = np.random.default_rng(2)
    p = rng.standard_normal((n_col, n_row)).astype(np.float32)
    pl = Pipeline(ctx, det)
    d = pl.load(p)
    pl.filter(d)
    got = pl.download(d)
    pl.release(d)
    pl.close()
    ref = port.filter(p, odet)
    scale = np.abs(ref).max()
    # float32 FFT against the oracle's float64 FFT
    assert np.abs(got - ref).max() <= 2e-6 * scale


def test_weight_filter_fused_equals_stages(ctx, port):
    odet, det = both_det(200, 50, l_px=0.3, delta_s=2.0)
    rng = np.random.default_rng(3)
    p = rng.standard_normal((50, 200)).astype(np.float32)
    pl = Pipeline(ctx, det)
    a = pl.load(p)
    pl.weight(a)
    pl.filter(a)
    b = pl.load(p)
    pl.weight_filter(b)
    ga, gb = pl.download(a), pl.download(b)
    pl.release(a)
    pl.release(b)
    pl.close()
    # the fused kernel computes the weight with one rsqrt per pixel (<= 2 ulp off the exact stand-alone kernel)
    assert np.abs(ga - gb).max() <= 1e-6 * np.abs(ga).max()
    ref = port.filter(port.weight(p, odet), odet)
    assert np.abs(ga - ref).max() <= 2e-6 * np.abs(ref).max()


@pytest.mark.parametrize("layout", [capi.LAYOUT_PLAIN, capi.LAYOUT_SPLIT2])
def test_filter_two_row_pairs_per_cta_equals_one(ctx, port, layout):
    """4096-point transforms run two row pairs per 512-thread CTA ("filter_wide", the default) or one per 256 threads:
    the same arithmetic per transform, so the same bits -- rows that do not fill the last 8-row block, both stack
    layouts, float and 16-bit samples, and the row-major form against the oracle."""
    n_row, n_col, n_proj = 2048, 43, 3
    odet, det = both_det(n_row, n_col, l_px=0.2, n_proj=n_proj)
    rng = np.random.default_rng(11)
    counts = rng.integers(0, 65536, size=(n_proj, n_col, n_row), dtype=np.uint16)
    as_float = counts.astype(np.float32)
    px = n_row * n_col
    d_f32, d_u16 = ctx.dev_alloc(as_float.nbytes), ctx.dev_alloc(counts.nbytes)
    ctx.vol_h2d(as_float, d_f32, as_float.size)
    ctx.vol_h2d(counts.view(np.float32), d_u16, counts.size // 2)
    filt = ctx.filter_create(capi.filter_size(n_row), float(det.l_px_row))
    slot_bytes, _ = capi.stack_slot_bytes(n_row, n_col)
    got = {}
    try:
        for wide in (1, 0):
            ctx.set_option("filter_wide", wide)
            for u16 in (False, True):
                stack = ctx.stack_alloc(n_row, n_col, n_proj)
                ctx.vol_h2d(np.zeros(n_proj * slot_bytes // 4, np.float32), stack, n_proj * slot_bytes // 4)
                if u16:
                    ctx.filter_to_stack_batch_u16(d_u16, px, n_proj, det, filt, stack, 0, layout)
                else:
                    ctx.filter_to_stack_batch(d_f32, px, n_proj, det, filt, stack, 0, layout)
                out = np.empty(n_proj * slot_bytes // 4, np.float32)
                ctx.vol_d2h(stack, out, out.size)
                ctx.stack_free(stack)
                got[(wide, u16)] = out
            # the row-major form (apply_filter of the contract) against the oracle
            pl = Pipeline(ctx, det)
            d = pl.load(as_float[0])
            pl.filter(d)
            rm = pl.download(d)
            pl.release(d)
            pl.close()
            ref = port.filter(as_float[0], odet)
            assert np.abs(rm - ref).max() <= 2e-6 * np.abs(ref).max()
            got[(wide, "rows")] = rm
    finally:
        ctx.set_option("filter_wide", 1)
        ctx.filter_destroy(filt)
        ctx.dev_free(d_f32)
        ctx.dev_free(d_u16)
    assert np.abs(got[(1, False)]).max() > 0
    assert np.array_equal(got[(1, False)], got[(0, False)]) and np.array_equal(got[(1, True)], got[(0, True)])
    assert np.array_equal(got[(1, False)], got[(1, True)])
    assert np.array_equal(got[(1, "rows")], got[(0, "rows")])


def _recon_case(n, n_proj, coarse=None, delta_s=0.0):
    odet, det = both_det(n, n, n_proj=n_proj, delta_s=delta_s)
    P = oracle.Port()
    ovol = P.calculate_volume_geometry(odet) if coarse is None else coarse_volume(odet, coarse)
    stack = shepp_logan(odet, n_proj)
    return odet, det, ovol, to_capi_vol(ovol), stack


@pytest.mark.parametrize("batch", [1, 5, 32])
def test_backproject_exact_kernel_bit_exact(ctx, port, batch):
    """bp_kernel=1: the reference's arithmetic operation for operation -> identical bits for any batch."""
    n, n_proj = 48, 12
    odet, det, ovol, vol, stack = _recon_case(n, n_proj)
    filtered = np.stack([port.filter(port.weight(s, odet), odet) for s in stack])
    ref = np.zeros((ovol.dim_z, ovol.dim_y, ovol.dim_x), np.float32)
    for i in range(n_proj):
        port.backproject(filtered[i], i, ref, odet, ovol)

    ctx.set_option("bp_kernel", 1)
    ctx.set_option("bp_batch", batch)
    pl = Pipeline(ctx, det)
    v = pl.make_volume(vol.dim_x, vol.dim_y, vol.dim_z)
    for i in range(n_proj):
        d = pl.load(filtered[i], idx=i)
        pl.backproject(d, v, 0, vol)
        pl.release(d)
    got = pl.save(v)
    pl.free_volume(v)
    pl.close()
    ctx.set_option("bp_kernel", 0)
    ctx.set_option("bp_batch", 64)
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("kernel", [1, 2])
def test_reconstruction_within_tolerance(ctx, port, kernel, fused):
    """config-1-like: K^3 volume from a (2K)^2 detector, full pipeline, north_star tolerance."""
    n, n_proj, k = 96, 64, 48
    odet, det, ovol, vol, stack = _recon_case(n, n_proj, coarse=k)
    ref, _ = port.reconstruct(stack, (k, k, k), odet, ovol)
    assert np.abs(ref).max() > 0.5 * contrast(n_proj)
    ctx.set_option("bp_kernel", kernel)
    pl = Pipeline(ctx, det)
    got = pl.reconstruct(stack, (k, k, k), vol, fused=fused)
    pl.close()
    ctx.set_option("bp_kernel", 0)
    mx, rms = errors(got, ref, contrast(n_proj))
    print(f"kernel={kernel} fused={fused}: max {mx:.3e} rmse {rms:.3e}")
    assert mx <= MAX_ABS_TOL and rms <= RMSE_TOL


def test_reupload_into_a_live_buffer_waits_for_the_kernels_queued_on_it(ctx, port):
    """A pooled device buffer that is uploaded to AGAIN (no dev_free / dev_alloc in between) while the fused
    weight+filter launch that reads its first content is still queued: the second upload must not overtake that
    launch (copies run on their own stream).  Slot 0 must hold filter(A), slot 1 filter(B)."""
    n_row, n_col = 1024, 256
    odet, det = both_det(n_row, n_col, l_px=0.2)
    rng = np.random.default_rng(7)
    a = rng.standard_normal((n_col, n_row)).astype(np.float32)
    b = rng.standard_normal((n_col, n_row)).astype(np.float32)
    ha, hb = capi.PinnedArray(a.shape), capi.PinnedArray(b.shape)
    ha.array[...] = a
    hb.array[...] = b
    f = ctx.filter_create(capi.filter_size(n_row), 0.2)
    slot_bytes, pitch = capi.stack_slot_bytes(n_row, n_col)
    st = ctx.stack_alloc(n_row, n_col, 2)
    d = ctx.dev_alloc(a.nbytes)
    for rep in range(20):   # (a race does not lose every time)
        ctx.proj_h2d(ha.ptr, d, n_row, n_col)
        ctx.filter_to_stack(d, det, f, st, 0, capi.LAYOUT_PLAIN)
        ctx.proj_h2d(hb.ptr, d, n_row, n_col)
        ctx.filter_to_stack(d, det, f, st, 1, capi.LAYOUT_PLAIN)
        out = np.empty((2, n_row, pitch), np.float32)
        ctx.proj_d2h(st, out[0], pitch, n_row)
        ctx.proj_d2h(st + slot_bytes, out[1], pitch, n_row)
        for got, src in ((out[0], a), (out[1], b)):
            ref = port.filter(port.weight(src, odet), odet)
            assert np.abs(got[:, :n_col].T - ref).max() <= 2e-6 * np.abs(ref).max(), rep
    ctx.dev_free(d)
    ctx.stack_free(st)
    ctx.filter_destroy(f)
    ha.free()
    hb.free()


def test_volume_alloc_lets_go_of_a_spare_slab_of_another_size(ctx):
    """volume_free keeps the last slab for the next volume_alloc of the SAME size; a request of another size (the
    last slab of a run carries the remainder) must release it first instead of holding two slabs."""
    a = ctx.volume_alloc(64, 64, 40)
    ctx.volume_free(a)                      # kept as the spare
    b = ctx.volume_alloc(64, 64, 43)        # another size: the spare goes
    c = ctx.volume_alloc(64, 64, 40)        # a fresh allocation, zero-initialised
    got = np.ones((40, 64, 64), np.float32)
    ctx.vol_d2h(c, got, got.size)
    assert not got.any()
    ctx.volume_free(b)
    ctx.volume_free(c)
    d = ctx.volume_alloc(64, 64, 40)        # the most recent spare (c) is reused
    assert d == c
    ctx.volume_free(d)


def test_volume_clear_drops_the_pending_batch(ctx, port):
    """volume_clear: 'pending batches targeting it are dropped' (include/paris_b200.h) -- projections enqueued before
    the clear must not show up in the volume, and the ones after it must."""
    n, n_proj = 48, 6
    odet, det, ovol, vol, stack = _recon_case(n, n_proj)
    pl = Pipeline(ctx, det)
    v = pl.make_volume(vol.dim_x, vol.dim_y, vol.dim_z)
    for i in range(3):
        d = pl.load(stack[i], idx=i)
        pl.backproject(d, v, 0, vol, fused_raw=True)
        pl.release(d)
    ctx.volume_clear(v.d_ptr, vol.dim_x, vol.dim_y, vol.dim_z)
    for i in range(3, n_proj):
        d = pl.load(stack[i], idx=i)
        pl.backproject(d, v, 0, vol, fused_raw=True)
        pl.release(d)
    got = pl.save(v)
    pl.free_volume(v)
    pl.close()
    ref = np.zeros((ovol.dim_z, ovol.dim_y, ovol.dim_x), np.float32)
    for i in range(3, n_proj):
        port.backproject(port.filter(port.weight(stack[i], odet), odet), i, ref, odet, ovol)
    mx, rms = errors(got, ref, contrast(n_proj))
    assert mx <= MAX_ABS_TOL and rms <= RMSE_TOL
