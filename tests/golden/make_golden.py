"""Regenerate tests/golden/*.npz from the REFERENCE ITSELF (oracle/_ref/libparis_ref.so, i.e. the unmodified
sources under /root/reference/src compiled by oracle/Makefile).  Run in the build container only:

    python tests/golden/make_golden.py

The fixtures pin the plain-C restatement (oracle/fdk_oracle.c) and, through it, the CUDA path, on boxes
where /root/reference does not exist.  Inputs are deterministic (analytic phantom, seeded noise)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
from cases import both_det, coarse_volume, shepp_logan  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (n_row, n_col, l_px, delta_s, delta_t, n_proj, coarse_k or None, roi or None)
    "natural_48x40": (48, 40, 0.4, 0.0, 0.0, 24, None, None),
    "offset_det_56x44": (56, 44, 0.3, 2.5, -1.5, 20, None, None),
    "coarse_64_to_32": (64, 64, 0.4, 0.0, 0.0, 32, 32, None),
    "roi_60x60": (60, 60, 0.4, 0.0, 0.0, 16, None, (10, 40, 8, 44, 5, 30)),
}


def main():
    for name, (n_row, n_col, l_px, ds, dt, n_proj, k, roi) in CASES.items():
        R = oracle.Reference()  # one private copy of the reference per geometry (function-local statics)
        odet, _ = both_det(n_row, n_col, l_px=l_px, delta_s=ds, delta_t=dt, n_proj=n_proj)
        ovol = R.calculate_volume_geometry(odet) if k is None else coarse_volume(odet, k)
        stack = shepp_logan(odet, n_proj)
        rng = np.random.default_rng(20261018)
        stack = (stack + rng.normal(0.0, 1e-3 * stack.max(), stack.shape)).astype(np.float32)
        weighted0 = R.weight(stack[0], odet)
        filtered0 = R.filter(weighted0, odet)
        if roi is None:
            r, shape = None, (ovol.dim_z, ovol.dim_y, ovol.dim_x)
        else:
            r = oracle.Roi(*roi)
            g = R.apply_roi(ovol, r)
            shape = (g.dim_z, g.dim_y, g.dim_x)
        vol, _ = R.reconstruct(stack, shape, odet, ovol, roi=r)
        kfilt = R.make_filter(2 * 2 ** int(np.ceil(np.log2(n_row))), l_px)
        np.savez_compressed(os.path.join(HERE, name + ".npz"),
                            det=np.array([n_row, n_col, l_px, l_px, ds, dt, 500.0, 500.0, 360.0 / n_proj], np.float64),
                            vol_geo=np.array([ovol.dim_x, ovol.dim_y, ovol.dim_z, ovol.l_vx_x, ovol.l_vx_y, ovol.l_vx_z],
                                             np.float64),
                            roi=np.array(roi if roi is not None else [], np.int64),
                            stack=stack.astype(np.float16) if False else stack,
                            weighted0=weighted0, filtered0=filtered0, k=kfilt, volume=vol)
        print(name, stack.shape, vol.shape, float(np.abs(vol).max()))


if __name__ == "__main__":
    main()
