"""Synthetic cone-beam projections of analytic ellipsoid phantoms (3-D Shepp-Logan).

Line integrals are evaluated in float64 on the reference's detector convention and
rounded once to float32, so the CPU oracle and the CUDA path can be fed identical inputs:

* rotation sense  s = x cos(phi) + y sin(phi),  t = -x sin(phi) + y cos(phi)
  (/root/reference/src/openmp/backprojection.cpp:121-122);
* source at s = -d_so, detector plane at s = +d_od, magnification d_sd / (s + d_so) (:125);
* pixel j of a detector row has its centre at h = -n_row*l/2 - delta_s*l + (j + 1/2)*l and row i
  at v = -n_col*l/2 - delta_t*l + (i + 1/2)*l (the backprojection's convention,
  src/openmp/backprojection.cpp:45-50).

TEST INFRASTRUCTURE (lives under oracle/): the numpy generator feeds the CPU checkers and the reference arm of
bench.py, which must not import the product.  The product's generator is the CUDA kernel csrc/phantom.cu (same
formulas; tests/test_gpu_cases.py::test_phantom_kernel_matches_numpy checks the two against each other, and
tests/test_oracle.py checks that the ellipsoid tables here and in paris_b200/phantom.py are the same).
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


def project(ellipsoids_mm: np.ndarray, n_row: int, n_col: int, l_px_row: float, l_px_col: float,
            delta_s: float, delta_t: float, d_so: float, d_od: float, angles_deg) -> np.ndarray:
    """Return the stack (n_proj, n_col, n_row) float32 of line integrals (density x millimetres)."""
    angles = np.atleast_1d(np.asarray(angles_deg, dtype=np.float64))
    d_sd = abs(d_so) + abs(d_od)
    h = -n_row * l_px_row / 2.0 - delta_s * l_px_row + (np.arange(n_row) + 0.5) * l_px_row
    v = -n_col * l_px_col / 2.0 - delta_t * l_px_col + (np.arange(n_col) + 0.5) * l_px_col
    hh, vv = np.meshgrid(h, v)  # (n_col, n_row)
    out = np.zeros((angles.size, n_col, n_row), dtype=np.float64)
    for ia, phi in enumerate(np.deg2rad(angles)):
        c, s_ = np.cos(phi), np.sin(phi)
        # source and detector point in world coordinates: x = s cos - t sin, y = s sin + t cos
        src = np.array([-d_so * c, -d_so * s_, 0.0])
        s_det = d_sd - d_so  # detector plane at s = +d_od (d_so > 0)
        dx = (s_det * c - hh * s_) - src[0]
        dy = (s_det * s_ + hh * c) - src[1]
        dz = vv - src[2]
        norm = np.sqrt(dx * dx + dy * dy + dz * dz)
        acc = np.zeros_like(hh)
        for rho, a, b, cax, x0, y0, z0, th in ellipsoids_mm:
            ct, st = np.cos(np.deg2rad(th)), np.sin(np.deg2rad(th))
            # into the ellipsoid frame (rotate by -theta about z), then scale to the unit sphere
            px, py, pz = src[0] - x0, src[1] - y0, src[2] - z0
            p0 = np.array([(px * ct + py * st) / a, (-px * st + py * ct) / b, pz / cax])
            d0 = (dx * ct + dy * st) / a
            d1 = (-dx * st + dy * ct) / b
            d2 = dz / cax
            A = d0 * d0 + d1 * d1 + d2 * d2
            B = p0[0] * d0 + p0[1] * d1 + p0[2] * d2
            Cc = p0 @ p0 - 1.0
            disc = B * B - A * Cc
            chord = np.where(disc > 0.0, 2.0 * np.sqrt(np.maximum(disc, 0.0)) / A, 0.0)
            acc += rho * chord * norm
        out[ia] = acc
    return out.astype(np.float32)


def shepp_logan_stack(n_row, n_col, l_px_row, l_px_col, delta_s, delta_t, d_so, d_od, n_proj,
                      delta_phi=None, fill=0.9, table=SHEPP_LOGAN_3D) -> np.ndarray:
    """Full-scan stack of the Shepp-Logan phantom scaled to fill*FOV radius; angle i = i*delta_phi."""
    if delta_phi is None:
        delta_phi = 360.0 / n_proj
    r = fill * fov_radius(n_row, l_px_row, delta_s, d_so, d_od)
    angles = np.float32(delta_phi) * np.arange(n_proj, dtype=np.float32)  # as the reference: float(idx)*delta_phi
    return project(scaled_ellipsoids(table, r), n_row, n_col, l_px_row, l_px_col, delta_s, delta_t,
                   d_so, d_od, angles.astype(np.float64))
