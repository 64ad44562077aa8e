import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import oracle
from paris_b200 import capi
from paris_b200.pipeline import Pipeline
from cases import both_det, shepp_logan, to_capi_vol, contrast, errors
n_row, n_col, n_proj = [int(x) for x in sys.argv[1:4]]
port = oracle.Port()
odet, det = both_det(n_row, n_col, n_proj=max(n_proj, 8))
ovol = port.calculate_volume_geometry(odet)
print("vol", ovol.dim_x, ovol.dim_y, ovol.dim_z, ovol.l_vx_x)
stack = shepp_logan(odet, n_proj)
shape = (ovol.dim_z, ovol.dim_y, ovol.dim_x)
ref, _ = port.reconstruct(stack, shape, odet, ovol)
ctx = capi.Context(0)
for kernel in (1, 2):
    for fused in (False, True):
        ctx.set_option("bp_kernel", kernel)
        pl = Pipeline(ctx, det)
        got = pl.reconstruct(stack, (ovol.dim_x, ovol.dim_y, ovol.dim_z), to_capi_vol(ovol), fused=fused)
        pl.close()
        d = np.abs(got.astype(np.float64) - ref)
        i = np.unravel_index(np.argmax(d), d.shape)
        print(f"kernel {kernel} fused {fused}: max {d.max()/contrast(max(n_proj,8)):.3e} at z,y,x={i} ref {ref[i]:.4f} got {got[i]:.4f}; bad voxels {(d > 1e-3).sum()}")
        if d.max() > 1e-3:
            bad = np.argwhere(d > 1e-3)
            print("  bad z range", bad[:,0].min(), bad[:,0].max(), "y", bad[:,1].min(), bad[:,1].max(), "x", bad[:,2].min(), bad[:,2].max())
