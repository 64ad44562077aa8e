"""BASELINE configurations 2-5 at FULL size against the CPU oracle.

No CPU implementation finishes 1e11 .. 2.5e13 voxel updates inside a test run, so the oracle is applied
block by block: the reference's ROI path (src/openmp/backprojection.cpp:105-118) reconstructs small boxes of voxels --
centre, an x-y edge, corners, slab seams -- from exactly the raw projections the GPU used, and each box is compared
with the crop of the GPU volume.  A box only ever reads a band of detector rows, so the oracle weights and filters
just that band (bit-identical to the same rows of the whole projection, tests/test_oracle_blocks.py); every other row
of its scratch projection is NaN.  The raw projections are generated on the device (float64 line integrals, the
kernel tests/test_gpu_cases.py checks against numpy) and the bands are downloaded from there: inputs are identical by
construction.

Besides the oracle blocks every configuration keeps its size-independent properties: the production kernel against
the exact kernel (pinned bit for bit to the reference on the small cases) on whole bands of slices, slabs and ROI
boxes as bit-identical crops, the phantom's plateau.  Every measured (max, rmse) / C goes to
gpurun_out/r2_parity.jsonl (copied to profiles/ by hand).
"""
import json
import os
import time

import numpy as np
import pytest

import oracle
from paris_b200 import capi, phantom
from paris_b200.pipeline import angle_sin_cos

from cases import MAX_ABS_TOL, RMSE_TOL, box_roi, contrast, errors, row_band, to_oracle_vol

pytestmark = pytest.mark.gpu

CHUNK = 64   # raw projections resident at a time


def _record(config, what, mx, rms, **extra):
    rec = {"config": config, "check": what, "max_over_C": mx, "rmse_over_C": rms, **extra}
    print(f"[parity] {config}: {what}: max {mx:.2e} C, rmse {rms:.2e} C")
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "r2_parity.jsonl"), "a") as f:
            f.write(json.dumps(rec) + "\n")
    except OSError:
        pass


class Scan:
    """One configuration on the device: the filtered stack built chunk by chunk, and the oracle's boxes built from
    the same raw projections on the way."""

    def __init__(self, ctx, port, name, det, vol_full, n_proj, blocks):
        self.ctx, self.port, self.name, self.det, self.vol, self.n_proj = ctx, port, name, det, vol_full, n_proj
        self.odet = oracle.DetectorGeometry(det.n_row, det.n_col, det.l_px_row, det.l_px_col, det.delta_s, det.delta_t,
                                            det.d_so, det.d_od, det.delta_phi)
        self.ovol = to_oracle_vol(vol_full)
        n = det.n_row
        self.px = n * det.n_col
        self.layout = capi.choose_stack_layout(det, vol_full)
        self.stack = ctx.stack_alloc(n, det.n_col, n_proj)
        self.filt = ctx.filter_create(capi.filter_size(n), float(det.l_px_row))
        self.sc = np.array([angle_sin_cos(i, det) for i in range(n_proj)], dtype=np.float32)
        self.c = contrast(n_proj)
        # blocks: name -> (x1, nx, y1, ny, z1, nz) in voxel indices of the FULL volume
        self.blocks = {}
        for bname, box in blocks.items():
            row0, n_rows = row_band(self.odet, self.ovol, *box)
            self.blocks[bname] = {"box": box, "row0": row0, "n_rows": n_rows, "roi": box_roi(*box), "vol": None}
        ell = phantom.scaled_ellipsoids(phantom.SHEPP_LOGAN_3D,
                                        0.9 * phantom.fov_radius(n, det.l_px_row, det.delta_s, det.d_so, det.d_od))
        raw = ctx.dev_alloc(CHUNK * self.px * 4)
        t_cpu = 0.0
        for first in range(0, n_proj, CHUNK):
            cnt = min(CHUNK, n_proj - first)
            ctx.phantom_project(ell, det, first, cnt, raw)
            for b in self.blocks.values():
                band = np.empty((cnt, b["n_rows"], n), np.float32)
                for i in range(cnt):
                    ctx.proj_d2h(raw + (i * self.px + b["row0"] * n) * 4, band[i], n, b["n_rows"])
                t0 = time.perf_counter()
                b["vol"] = port.reconstruct_block(band, b["row0"], self.odet, self.ovol, b["roi"], first_idx=first,
                                                  vol=b["vol"])
                t_cpu += time.perf_counter() - t0
            ctx.filter_to_stack_batch(raw, self.px, cnt, det, self.filt, self.stack, first, self.layout)
        ctx.dev_free(raw)
        print(f"[parity] {name}: oracle boxes took {t_cpu:.1f} s of CPU "
              f"(bands of {[b['n_rows'] for b in self.blocks.values()]} rows)")

    def backproject(self, dims, v_offset, roi=None, kernel=2, d_vol=None):
        """(dims[2], dims[1], dims[0]) float32 from the production (2) or exact (1) kernel, and the milliseconds"""
        ctx = self.ctx
        ctx.set_option("bp_kernel", kernel)
        own = d_vol is None
        if own:
            d_vol = ctx.volume_alloc(*dims)
        e0 = ctx.event()
        ctx.backproject_stack(self.stack, 0, self.n_proj, self.sc[:, 0], self.sc[:, 1], d_vol, dims, v_offset, self.det,
                              self.vol, roi=roi, layout=self.layout)
        e1 = ctx.event()
        ms = ctx.elapsed_ms(e0, e1)
        out = None
        if own:
            out = np.empty((dims[2], dims[1], dims[0]), np.float32)
            ctx.vol_d2h(d_vol, out, out.size)
            ctx.volume_free(d_vol)
        ctx.set_option("bp_kernel", 0)
        return out, ms

    def check_blocks(self, volume, origin):
        """volume: the GPU region (dz, dy, dx) whose voxel (0, 0, 0) is voxel `origin` = (x, y, z) of the full volume"""
        worst = (0.0, 0.0)
        for bname, b in self.blocks.items():
            x1, nx, y1, ny, z1, nz = b["box"]
            zo, yo, xo = z1 - origin[2], y1 - origin[1], x1 - origin[0]
            if zo < 0 or zo + nz > volume.shape[0]:
                continue   # (a box outside the part of the region that was reconstructed)
            got = volume[zo:zo + nz, yo:yo + ny, xo:xo + nx]
            assert got.shape == b["vol"].shape, (bname, got.shape, b["vol"].shape)
            assert np.isfinite(b["vol"]).all(), f"{bname}: the oracle read outside its band of rows"
            peak = float(np.abs(b["vol"]).max())
            if peak == 0.0:
                # detector rows that see nothing of the phantom: the reference adds exact zeros, so must the kernel
                assert not got.any(), f"{bname}: the reference leaves this box untouched"
            mx, rms = errors(got, b["vol"], self.c)
            _record(self.name, f"oracle ROI block '{bname}' {b['box']}", mx, rms, band_rows=b["n_rows"],
                    oracle_peak_over_C=peak / self.c)
            assert mx <= MAX_ABS_TOL and rms <= RMSE_TOL, (bname, mx, rms)
            worst = (max(worst[0], mx), max(worst[1], rms))
        return worst

    def close(self):
        self.ctx.filter_destroy(self.filt)
        self.ctx.stack_free(self.stack)


def _coarse(n, n_proj, k, l_px):
    det = capi.DetectorGeometry(n, n, l_px, l_px, 0, 0, 500, 500, 360.0 / n_proj)
    nat = capi.calculate_volume_geometry(det)
    f32 = np.float32
    vol = capi.VolumeGeometry(k, k, k, f32(nat.l_vx_x * nat.dim_x / k), f32(nat.l_vx_y * nat.dim_y / k),
                              f32(nat.l_vx_z * nat.dim_z / k))
    return det, vol


def test_config2_full_size_against_oracle_blocks(ctx, port):
    """512^3 from 720 x 1024^2 (coarse volume, split stack layout)."""
    n, n_proj, k = 1024, 720, 512
    det, vol = _coarse(n, n_proj, k, 0.2)
    # (the phantom ends at 0.73 of the volume's half height; slices much beyond that only see empty detector rows)
    blocks = {"centre": (240, 32, 240, 32, 240, 32), "x-y edge": (0, 32, 240, 32, 240, 32),
              "x-y corner": (480, 32, 480, 32, 240, 32), "high z, |v| ~ 400..470 rows": (240, 32, 240, 32, 440, 32),
              "low z at the x-y edge": (0, 32, 240, 32, 40, 32), "top corner (empty rows)": (480, 32, 480, 32, 480, 32),
              "seam z=256": (150, 16, 350, 16, 252, 8)}
    s = Scan(ctx, port, "c2", det, vol, n_proj, blocks)
    assert s.layout == capi.LAYOUT_SPLIT2
    fast, ms = s.backproject((k, k, k), 0)
    assert "bp_tma_kernel<8x8x128" in ctx.bp_kernel_info()["last"]
    print(f"config 2 full size: {k ** 3 * n_proj / ms / 1e6:.0f} GUPS ({ms:.1f} ms)")
    s.check_blocks(fast, (0, 0, 0))
    exact, _ = s.backproject((k, k, k), 0, kernel=1)
    s.close()
    mx, rms = errors(fast, exact, s.c)
    _record("c2", "production kernel vs exact kernel, whole volume", mx, rms)
    assert mx <= MAX_ABS_TOL and rms <= RMSE_TOL
    # brain tissue of the phantom (density 1.0 - 0.8 = 0.2) a few voxels off the centre, away from the ventricles
    plateau = fast[k // 2, k // 2 - 40:k // 2 - 30, k // 2 - 4:k // 2 + 4].mean()
    assert abs(plateau / (0.2 * s.c) - 1.0) < 0.05


def test_config3_full_size_against_oracle_blocks_and_slab_seams(ctx, port):
    """The north-star configuration: 1024^3 from 1440 x 2048^2, reconstructed as the eight 128-slice slabs the
    8-GPU run cuts it into (same launches: dims (1024, 1024, 128), v_offset 128 s).  Oracle boxes at the centre, an
    x-y edge and corner, high and low slices (|v| ~ 690..830 rows from the detector's centre, where the reference's
    float32 row rounding is coarsest; beyond 0.73 of the half height the rows see nothing of the phantom) and across
    all seven slab seams; the production kernel against the exact kernel on bands at the bottom, middle and
    top and across every seam (z = 128 s +- 2)."""
    n, n_proj, k, slab = 2048, 1440, 1024, 128
    det, vol = _coarse(n, n_proj, k, 0.1)
    blocks = {"centre": (496, 32, 496, 32, 496, 32), "x-y edge": (0, 32, 496, 32, 496, 32),
              "x-y corner": (992, 32, 992, 32, 496, 32), "high z, |v| ~ 690..830 rows": (496, 32, 496, 32, 856, 32),
              "low z at the x-y edge": (0, 32, 496, 32, 96, 32), "bottom corner (empty rows)": (0, 32, 0, 32, 0, 32)}
    for sidx in range(1, k // slab):
        blocks[f"seam z={slab * sidx}"] = (300, 16, 700, 16, slab * sidx - 4, 8)
    s = Scan(ctx, port, "c3", det, vol, n_proj, blocks)
    assert s.layout == capi.LAYOUT_SPLIT2
    d_vol = ctx.volume_alloc(k, k, k)
    total = 0.0
    exact_before = ctx.bp_kernel_info()["exact_launches"]
    for sidx in range(k // slab):
        _, ms = s.backproject((k, k, slab), sidx * slab, d_vol=d_vol + sidx * slab * k * k * 4)
        total += ms
    info = ctx.bp_kernel_info()
    assert "bp_tma_kernel<8x8x128" in info["last"] and info["exact_launches"] == exact_before, info
    print(f"config 3 full size, 8 slabs: {k ** 3 * n_proj / total / 1e6:.0f} GUPS ({total:.0f} ms)")
    fast = np.empty((k, k, k), np.float32)
    ctx.vol_d2h(d_vol, fast, fast.size)
    ctx.volume_free(d_vol)
    s.check_blocks(fast, (0, 0, 0))
    band = 4
    zs = [0, k // 2 - 2, k - band] + [slab * i - 2 for i in range(1, k // slab)]
    for z in zs:
        exact, _ = s.backproject((k, k, band), z, kernel=1)
        mx, rms = errors(fast[z:z + band], exact, s.c)
        _record("c3", f"production kernel (8 slabs) vs exact kernel, slices [{z}, {z + band})", mx, rms)
        assert mx <= MAX_ABS_TOL and rms <= RMSE_TOL
    # one piece == eight slabs, bit for bit (two slabs' worth is enough to cross a seam)
    piece, _ = s.backproject((k, k, 2 * slab), 3 * slab)
    assert np.array_equal(piece, fast[3 * slab:5 * slab])
    s.close()
    plateau = fast[k // 2, k // 2 - 80:k // 2 - 60, k // 2 - 8:k // 2 + 8].mean()
    assert abs(plateau / (0.2 * s.c) - 1.0) < 0.05


def test_config4_roi_offset_detector_against_oracle_blocks(ctx, port):
    """1024^3 region of interest of the 2248 x 2248 x 2060 natural volume from 2880 projections of a 2048^2 detector
    shifted by 100 pixels (plain stack layout, STRADDLE tiles: the ROI's z offset 518 is no multiple of the tile)."""
    n, n_proj, k = 2048, 2880, 1024
    l_px = 0.1
    det = capi.DetectorGeometry(n, n, l_px, l_px, 100.0, 0, 500, 500, 360.0 / n_proj)
    nat = capi.calculate_volume_geometry(det)
    assert (nat.dim_x, nat.dim_y, nat.dim_z) == (2248, 2248, 2060)
    roi = capi.Roi(612, 1636, 612, 1636, 518, 1542)
    reg = capi.apply_roi(nat, roi)
    assert (reg.dim_x, reg.dim_y, reg.dim_z) == (k, k, k)
    blocks = {"centre": (1108, 32, 1108, 32, 1014, 32), "ROI x-y edge": (612, 32, 1108, 32, 1014, 32),
              "ROI top corner": (1604, 32, 1604, 32, 1510, 32), "ROI bottom corner": (612, 32, 612, 32, 518, 32),
              "row-anchor and tile seam": (900, 16, 1300, 16, 518 + 118, 14)}   # ROI z 118..132: anchor at 122, tiles at 128
    s = Scan(ctx, port, "c4", det, nat, n_proj, blocks)
    assert s.layout == capi.LAYOUT_PLAIN
    full, ms = s.backproject((k, k, k), 0, roi=roi)
    print(f"config 4 full size: {k ** 3 * n_proj / ms / 1e6:.0f} GUPS ({ms:.0f} ms)")
    s.check_blocks(full, (612, 612, 518))
    band = 4
    for z in (0, 509, k - band):                                        # bands of `band` slices at these ROI offsets
        fast_band, _ = s.backproject((k, k, band), z, roi=roi)
        assert np.array_equal(fast_band, full[z:z + band]), f"band at {z} is not a crop of the one-piece ROI"
        exact, _ = s.backproject((k, k, band), z, roi=roi, kernel=1)
        mx, rms = errors(fast_band, exact, s.c)
        _record("c4", f"production kernel vs exact kernel, ROI slices [{z}, {z + band})", mx, rms)
        assert mx <= MAX_ABS_TOL and rms <= RMSE_TOL
    s.close()
    assert np.isfinite(full).all() and full.max() > 0.5 * 1.0 * s.c * 0.2


def test_config5_slabs_streamed_over_one_stack_against_oracle_blocks(ctx, port):
    """The 2048^3 region (z in [5, 2053) of the 2048 x 2048 x 2058 natural volume) from 2880 projections of a
    2048^2 detector, reconstructed slab by slab from ONE filtered stack (the reference re-reads and re-filters the
    whole scan per sub-volume, src/main.cpp:93-105).  Two adjacent 128-slice slabs of the 16: streamed separately they
    are bit-identical to the same 256 slices in one piece; oracle boxes inside, across their seam and at the slabs'
    x-y corners."""
    n, n_proj, k, slab = 2048, 2880, 2048, 128
    l_px = 0.1
    det = capi.DetectorGeometry(n, n, l_px, l_px, 0, 0, 500, 500, 360.0 / n_proj)
    nat = capi.calculate_volume_geometry(det)
    assert (nat.dim_x, nat.dim_y, nat.dim_z) == (2048, 2048, 2058)
    roi = capi.Roi(0, 2047, 0, 2047, 5, 2053)
    reg = capi.apply_roi(nat, roi)
    assert (reg.dim_x, reg.dim_y, reg.dim_z) == (k, k, k)
    z0 = 5 + 7 * slab
    blocks = {"inside slab 7": (1008, 32, 1008, 32, z0 + 48, 32), "seam of slabs 7|8": (500, 16, 1500, 16, z0 + slab - 4, 8),
              "x-y corner (0, 0)": (0, 32, 0, 32, z0, 16), "x-y corner (2047, 2047)": (2016, 32, 2016, 32, z0 + 2 * slab - 16, 16)}
    s = Scan(ctx, port, "c5", det, nat, n_proj, blocks)
    a, ms_a = s.backproject((k, k, slab), 7 * slab, roi=roi)
    b, _ = s.backproject((k, k, slab), 8 * slab, roi=roi)
    print(f"config 5 slab of {slab} slices: {k * k * slab * n_proj / ms_a / 1e6:.0f} GUPS ({ms_a:.0f} ms)")
    both, _ = s.backproject((k, k, 2 * slab), 7 * slab, roi=roi)
    assert np.array_equal(both[:slab], a) and np.array_equal(both[slab:], b)
    s.check_blocks(both, (0, 0, z0))
    for z in (7 * slab, 9 * slab - 2):
        exact, _ = s.backproject((k, k, 2), z, roi=roi, kernel=1)
        mx, rms = errors(both[z - 7 * slab:z - 7 * slab + 2], exact, s.c)
        _record("c5", f"production kernel vs exact kernel, ROI slices [{z}, {z + 2})", mx, rms)
        assert mx <= MAX_ABS_TOL and rms <= RMSE_TOL
    s.close()


def test_row_rounding_error_does_not_grow_with_the_detector(ctx, port):
    """How the production kernel's distance from the reference's arithmetic scales with detector size.  The kernel
    carries detector rows in fixed point (2^-22 rows); the reference rounds v to float32, whose ulp DOUBLES with every
    doubling of the detector (6e-5 rows at |v| ~ 1000, 1.2e-4 at ~ 2000) -- but the same object edge then spans twice
    as many rows, so the filtered projection changes half as much per row and the product stays put.  Measured here on
    one 4-slice band at 0.7 of the half height (where the phantom still has structure) of K^3 volumes from (2K)^2
    detectors, K = 256 .. 1024, 360 projections each, production kernel against the exact kernel.

    History: round 1 carried detector rows in fixed point -- more accurately than the reference, which rounds v to
    float32 after every operation (ulp 6e-5 rows at |v| ~ 1000, 1.2e-4 at ~ 2000) -- and was 6e-5 .. 1.05e-4 C away
    from it on single voxels at every detector size (the ulp doubles with the detector, the slope of the filtered
    projection per row halves).  The kernel now evaluates the row in the reference's own float operations and sits at
    ~3e-7 C whatever the size."""
    n_proj = 360
    results = {}
    for n in (512, 1024, 2048):
        k = n // 2
        det, vol = _coarse(n, n_proj, k, 0.2 * 1024 / n)
        s = Scan(ctx, port, f"growth-{n}", det, vol, n_proj, {})
        z = int(k // 2 + 0.7 * (k // 2))
        fast, _ = s.backproject((k, k, 4), z)
        exact, _ = s.backproject((k, k, 4), z, kernel=1)
        s.close()
        assert np.abs(exact).max() > 0.05 * s.c
        mx, rms = errors(fast, exact, s.c)
        _record(f"growth-{n}", f"production vs exact kernel, {k}^3-equivalent band [{z}, {z + 4}) from {n_proj} x {n}^2", mx, rms)
        results[n] = mx
        assert mx <= MAX_ABS_TOL and rms <= RMSE_TOL
    # no growth: the largest detector is within 2x of the smallest (4x if the error followed the reference's ulp)
    assert results[2048] < 2.0 * max(results[512], results[1024])
    assert max(results.values()) < 2e-6
