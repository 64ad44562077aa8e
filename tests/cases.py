"""Shared geometry / phantom cases for the parity tests."""
from __future__ import annotations

import numpy as np

import oracle
from paris_b200 import capi, phantom


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
