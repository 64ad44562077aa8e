"""Shared geometry / phantom cases for the parity tests."""
from __future__ import annotations

import numpy as np

import oracle
import oracle.phantom as phantom
from paris_b200 import capi


def both_det(n_row, n_col, l_px=0.4, delta_s=0.0, delta_t=0.0, d_so=500.0, d_od=500.0, n_proj=64, l_px_col=None):
    """The same detector geometry as the oracle's and the product's ctypes struct."""
    args = (n_row, n_col, l_px, l_px if l_px_col is None else l_px_col, delta_s, delta_t, d_so, d_od, 360.0 / n_proj)
    return oracle.DetectorGeometry(*args), capi.DetectorGeometry(*args)


def to_capi_vol(v) -> capi.VolumeGeometry:
    return capi.VolumeGeometry(v.dim_x, v.dim_y, v.dim_z, v.l_vx_x, v.l_vx_y, v.l_vx_z)


def to_oracle_vol(v) -> oracle.VolumeGeometry:
    return oracle.VolumeGeometry(v.dim_x, v.dim_y, v.dim_z, v.l_vx_x, v.l_vx_y, v.l_vx_z)


def coarse_volume(det, k: int) -> oracle.VolumeGeometry:
    """'K^3 from a (2K)^2 detector' (BASELINE configs 1-3): the natural full-FOV volume of
    calculate_volume_geometry sampled with k^3 voxels of proportionally larger size."""
    nat = oracle.Port().calculate_volume_geometry(det)
    return oracle.VolumeGeometry(k, k, k,
                                 np.float32(nat.l_vx_x * nat.dim_x / k),
                                 np.float32(nat.l_vx_y * nat.dim_y / k),
                                 np.float32(nat.l_vx_z * nat.dim_z / k))


def shepp_logan(det, n_proj):
    return phantom.shepp_logan_stack(det.n_row, det.n_col, det.l_px_row, det.l_px_col, det.delta_s, det.delta_t,
                                     det.d_so, det.d_od, n_proj, delta_phi=det.delta_phi)


def contrast(n_proj: int) -> float:
    """Phantom contrast in reference output units: C = delta_mu * N_proj / (8 pi)."""
    return phantom.SHEPP_LOGAN_CONTRAST * n_proj / (8.0 * np.pi)


def errors(a: np.ndarray, b: np.ndarray, c: float):
    d = a.astype(np.float64) - b.astype(np.float64)
    return float(np.abs(d).max() / c), float(np.sqrt(np.mean(d * d)) / c)


# north_star tolerance: max abs error <= 1e-4 of phantom contrast, RMSE <= 1e-5
MAX_ABS_TOL = 1e-4
RMSE_TOL = 1e-5


# ---- full-size checks block by block (no CPU oracle finishes 1e12 updates) -----------------------------

def box_roi(x1, nx, y1, ny, z1, nz) -> "oracle.Roi":
    """The region_of_interest whose box is [x1, x1+nx) x [y1, y1+ny) x [z1, z1+nz) under apply_roi's rule
    (dim = x2 - x1, + 1 iff x1 == 0; /root/reference/src/geometry.cpp:86-130)."""
    def upper(lo, n):
        return lo + n - 1 if lo == 0 else lo + n
    return oracle.Roi(x1, upper(x1, nx), y1, upper(y1, ny), z1, upper(z1, nz))


def row_band(det, vol_full, x1, nx, y1, ny, z1, nz, margin=3):
    """[row0, row0 + n_rows): a conservative band of detector rows the voxels [x1, x1+nx) x [y1, y1+ny) x [z1, z1+nz)
    of the FULL volume can read at any projection angle: v = (z_m * factor - min_v) / l_px - 0.5 with
    factor = d_sd / (s + d_so), |s| <= the box's largest distance from the rotation axis
    (/root/reference/src/openmp/backprojection.cpp:120-133).  Float64; `margin` rows either side cover float32
    rounding of the reference's own v and the + 1 neighbour."""
    def centred(i, dim, size):
        return -(dim * size / 2.0) + size / 2.0 + i * size
    xs = [centred(i, vol_full.dim_x, float(vol_full.l_vx_x)) for i in (x1, x1 + nx - 1)]
    ys = [centred(i, vol_full.dim_y, float(vol_full.l_vx_y)) for i in (y1, y1 + ny - 1)]
    zs = [centred(i, vol_full.dim_z, float(vol_full.l_vx_z)) for i in (z1, z1 + nz - 1)]
    r = max(np.hypot(x, y) for x in xs for y in ys)
    d_so, d_sd = float(det.d_so), abs(float(det.d_so)) + abs(float(det.d_od))
    assert d_so - r > 0
    factors = (d_sd / (d_so + r), d_sd / (d_so - r))
    l_px = float(det.l_px_col)
    min_v = -(det.n_col * l_px / 2.0) - float(det.delta_t) * l_px
    vs = [(z * f - min_v) / l_px - 0.5 for z in zs for f in factors]
    lo = int(np.floor(min(vs))) - margin
    hi = int(np.floor(max(vs))) + 1 + margin
    lo, hi = max(lo, 0), min(hi, det.n_col - 1)
    if hi < lo:          # the box never sees the detector: any one row will do
        lo = hi = min(max(lo, 0), det.n_col - 1)
    return lo, hi - lo + 1
