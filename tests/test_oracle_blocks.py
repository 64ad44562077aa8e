"""The block-wise form of the oracle used by the full-size GPU checks (tests/test_gpu_fullsize.py) is the same
arithmetic as the whole-projection form: bands of detector rows filtered on their own and boxes reconstructed
through the reference's ROI path are bit-identical crops."""
import numpy as np
import pytest

import oracle

from cases import both_det, box_roi, coarse_volume, row_band, shepp_logan


def test_band_filter_is_a_bit_identical_crop(port):
    odet, _ = both_det(200, 90, l_px=0.3, delta_s=2.0, delta_t=-1.0)
    rng = np.random.default_rng(3)
    p = rng.standard_normal((90, 200)).astype(np.float32)
    full = port.filter(port.weight(p, odet), odet)
    for row0, n in ((0, 90), (0, 7), (31, 20), (89, 1)):
        band = port.weight_filter_rows(p, odet, row0, n)
        assert np.array_equal(band[row0:row0 + n], full[row0:row0 + n])
        keep = np.ones(90, bool)
        keep[row0:row0 + n] = False
        assert np.array_equal(band[keep], p[keep])


@pytest.mark.skipif(not oracle.have_ref(), reason="reference build absent")
def test_band_filter_equals_reference_at_full_detector_size(port):
    """one 2048^2 projection (configs 3-5): the band the restatement filters equals the reference's own rows"""
    n = 2048
    odet, _ = both_det(n, n, l_px=0.1, n_proj=1440)
    rng = np.random.default_rng(4)
    p = rng.standard_normal((n, n)).astype(np.float32)
    ref = oracle.Reference()
    full = ref.filter(ref.weight(p, odet), odet)
    band = port.weight_filter_rows(p, odet, 1650, 398)
    assert np.array_equal(band[1650:2048], full[1650:2048])


@pytest.mark.parametrize("coarse", [False, True])
def test_block_reconstruction_is_a_bit_identical_crop(port, coarse):
    """reconstruct_block (band of rows, NaN elsewhere, projections in two chunks) == crop of the whole reconstruction"""
    n, n_proj = 64, 24
    odet, _ = both_det(n, 72, n_proj=n_proj, delta_s=1.0)
    ovol = coarse_volume(odet, 32) if coarse else port.calculate_volume_geometry(odet)
    stack = shepp_logan(odet, n_proj)
    whole, _ = port.reconstruct(stack, (ovol.dim_z, ovol.dim_y, ovol.dim_x), odet, ovol)
    for (x1, nx, y1, ny, z1, nz) in ((0, 8, 5, 6, 0, 4), (ovol.dim_x - 9, 9, 10, 3, ovol.dim_z - 5, 5),
                                     (12, 4, 12, 4, ovol.dim_z // 2 - 2, 4)):
        roi = box_roi(x1, nx, y1, ny, z1, nz)
        row0, n_rows = row_band(odet, ovol, x1, nx, y1, ny, z1, nz)
        band = np.ascontiguousarray(stack[:, row0:row0 + n_rows, :])
        got = port.reconstruct_block(band[:10], row0, odet, ovol, roi)
        got = port.reconstruct_block(band[10:], row0, odet, ovol, roi, first_idx=10, vol=got)
        assert np.isfinite(got).all()
        assert np.array_equal(got, whole[z1:z1 + nz, y1:y1 + ny, x1:x1 + nx])


def test_a_band_chosen_too_small_is_loud(port):
    n, n_proj = 64, 4
    odet, _ = both_det(n, 72, n_proj=n_proj)
    ovol = port.calculate_volume_geometry(odet)
    stack = shepp_logan(odet, n_proj)
    roi = box_roi(20, 8, 20, 8, 30, 8)
    row0, n_rows = row_band(odet, ovol, 20, 8, 20, 8, 30, 8)
    bad = np.ascontiguousarray(stack[:, row0 + 4:row0 + n_rows - 4, :])
    got = port.reconstruct_block(bad, row0 + 4, odet, ovol, roi)
    assert not np.isfinite(got).all()
