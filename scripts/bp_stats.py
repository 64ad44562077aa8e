"""Developer probe (needs `make lib EXTRA_NVFLAGS=-DPB_BP_STATS`): how many tile-projections are interior / mixed /
skipped, and how many columns take the careful path, for a bench configuration."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from paris_b200 import capi
from paris_b200.pipeline import angle_sin_cos
import bench
cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
natural = len(sys.argv) > 2
det, vol, n_proj = bench.geometry(cfg)
if natural:
    vol = capi.calculate_volume_geometry(det)
ctx = capi.Context(0)
n = 64
slot_bytes, _ = capi.stack_slot_bytes(det.n_row, det.n_col)
stack = ctx.stack_alloc(det.n_row, det.n_col, n)
sc = np.array([angle_sin_cos(i * (n_proj // n), det) for i in range(n)], dtype=np.float32)
dims = (vol.dim_x, vol.dim_y, vol.dim_z)
v = ctx.volume_alloc(*dims)
layout = capi.choose_stack_layout(det, vol)
ctx.set_option("bp_kernel", 2)
ctx.backproject_stack(stack, 0, n, sc[:, 0], sc[:, 1], v, dims, 0, det, vol, layout=layout)
out = (C.c_ulonglong * 8)()
capi.lib().paris_b200_debug_bp_stats(out)
o = list(out)
tp = o[0] + o[1] + o[2]
print(f"{cfg} natural={natural} layout={layout}: tile-projections {tp}: mixed {o[0]/tp:.1%} interior {o[1]/tp:.1%} skipped {o[2]/tp:.1%}")
print(f"  table entries {o[3]}: careful {o[4]/max(o[3],1):.1%} dead {o[5]/max(o[3],1):.1%} in mixed tiles {o[6]/max(o[3],1):.1%}")
