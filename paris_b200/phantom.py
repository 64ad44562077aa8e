"""Synthetic cone-beam projections of analytic ellipsoid phantoms (3-D Shepp-Logan).

Line integrals are evaluated in float64 on the reference's detector convention and
rounded once to float32, so the CPU oracle and the CUDA path can be fed identical inputs:

* rotation sense  s = x cos(phi) + y sin(phi),  t = -x sin(phi) + y cos(phi)
  (/root/reference/src/openmp/backprojection.cpp:121-122);
* source at s = -d_so, detector plane at s = +d_od, magnification d_sd / (s + d_so) (:125);
* pixel j of a detector row has its centre at h = -n_row*l/2 - delta_s*l + (j + 1/2)*l and row i
  at v = -n_col*l/2 - delta_t*l + (i + 1/2)*l (the backprojection's convention,
  src/openmp/backprojection.cpp:45-50).

This module only holds the ellipsoid tables and their scaling; the projector itself is the CUDA kernel
csrc/phantom.cu (paris_b200_phantom_project).  The numpy projector that checks it lives with the other CPU
checkers in oracle/phantom.py.
"""
from __future__ import annotations

import numpy as np

# density, semi-axes a b c, centre x0 y0 z0, rotation about z (degrees); unit-ball coordinates.
# The high-contrast ("modified") 3-D Shepp-Logan table.
SHEPP_LOGAN_3D = np.array([
    [1.0, .6900, .920, .810, 0.0, 0.0, 0.0, 0.0],
    [-.8, .6624, .874, .780, 0.0, -.0184, 0.0, 0.0],
    [-.2, .1100, .310, .220, .22, 0.0, 0.0, -18.0],
    [-.2, .1600, .410, .280, -.22, 0.0, 0.0, 18.0],
    [.1, .2100, .250, .410, 0.0, .35, -.15, 0.0],
    [.1, .0460, .046, .050, 0.0, .1, .25, 0.0],
    [.1, .0460, .046, .050, 0.0, -.1, .25, 0.0],
    [.1, .0460, .023, .050, -.08, -.605, 0.0, 0.0],
    [.1, .0230, .023, .020, 0.0, -.606, 0.0, 0.0],
    [.1, .0230, .046, .020, .06, -.605, 0.0, 0.0],
], dtype=np.float64)

#: max - min density of SHEPP_LOGAN_3D (skull 1.0 against air 0.0)
SHEPP_LOGAN_CONTRAST = 1.0

UNIT_SPHERE = np.array([[1.0, 1.0, 1.0, 1.0, 0.0, 0.0, 0.0, 0.0]], dtype=np.float64)


def fov_radius(n_row, l_px_row, delta_s, d_so, d_od) -> float:
    """Radius of the cylinder every ray fan covers (the r of src/geometry.cpp:52-53)."""
    d_sd = abs(d_so) + abs(d_od)
    alpha = np.arctan((n_row * l_px_row / 2.0 + abs(delta_s * l_px_row)) / d_sd)
    return float(abs(d_so) * np.sin(alpha))


def scaled_ellipsoids(table: np.ndarray, radius_mm: float) -> np.ndarray:
    """Scale a unit-ball ellipsoid table to millimetres."""
    e = np.array(table, dtype=np.float64, copy=True)
    e[:, 1:7] *= radius_mm
    return e


def line_integral_bound(ellipsoids_mm: np.ndarray) -> float:
    """No ray through the phantom integrates to more than this: every ellipsoid of positive density at its longest
    chord.  counts_scale(...) turns it into the factor that maps line integrals onto 16-bit detector counts."""
    e = np.asarray(ellipsoids_mm, dtype=np.float64)
    pos = e[e[:, 0] > 0]
    return float((pos[:, 0] * 2.0 * pos[:, 1:4].max(axis=1)).sum())


def counts_scale(ellipsoids_mm: np.ndarray, full_scale: float = 60000.0) -> float:
    return full_scale / line_integral_bound(ellipsoids_mm)
