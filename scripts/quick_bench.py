"""Developer micro-bench: device-resident filter + backprojection timing for one config."""
import argparse
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from paris_b200 import capi, phantom
from paris_b200.pipeline import angle_sin_cos

ap = argparse.ArgumentParser()
ap.add_argument("--det", type=int, default=1024)
ap.add_argument("--vol", type=int, default=512)
ap.add_argument("--proj", type=int, default=720)
ap.add_argument("--natural", action="store_true")
ap.add_argument("--kernel", type=int, default=0)
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--layout", type=int, default=-1)
ap.add_argument("--tile", type=int, default=0)
ap.add_argument("--swizzle", type=int, default=16)
ap.add_argument("--check", action="store_true", help="compare against the exact kernel")
a = ap.parse_args()

l_px = 0.2 * 1024 / a.det
det = capi.DetectorGeometry(a.det, a.det, l_px, l_px, 0, 0, 500, 500, 360.0 / a.proj)
nat = capi.calculate_volume_geometry(det)
if a.natural:
    vol = nat
else:
    k = a.vol
    vol = capi.VolumeGeometry(k, k, k, nat.l_vx_x * nat.dim_x / k, nat.l_vx_y * nat.dim_y / k, nat.l_vx_z * nat.dim_z / k)
print("volume", vol.dim_x, vol.dim_y, vol.dim_z, vol.l_vx_x)
ctx = capi.Context(0)
ctx.set_option("bp_batch", a.batch)
ctx.set_option("bp_kernel", a.kernel)
ctx.set_option("bp_tile", a.tile)
ctx.set_option("bp_swizzle", a.swizzle)
n = a.proj
raw = ctx.dev_alloc(n * a.det * a.det * 4)
r = 0.9 * phantom.fov_radius(a.det, l_px, 0, 500, 500)
ctx.phantom_project(phantom.scaled_ellipsoids(phantom.SHEPP_LOGAN_3D, r), det, 0, n, raw)
slot_bytes, pitch = capi.stack_slot_bytes(a.det, a.det)
stack = ctx.stack_alloc(a.det, a.det, n)
filt = ctx.filter_create(capi.filter_size(a.det), l_px)
sc = np.array([angle_sin_cos(i, det) for i in range(n)], dtype=np.float32)
dims = (vol.dim_x, vol.dim_y, vol.dim_z)
layout = capi.choose_stack_layout(det, vol) if a.layout < 0 else a.layout
print('layout', layout)
d_vol = ctx.volume_alloc(*dims)
updates = vol.dim_x * vol.dim_y * vol.dim_z * n
for rep in range(a.reps):
    e0 = ctx.event()
    ctx.filter_to_stack_batch(raw, a.det * a.det, n, det, filt, stack, 0, layout)
    e1 = ctx.event()
    ctx.backproject_stack(stack, 0, n, sc[:, 0], sc[:, 1], d_vol, dims, 0, det, vol, layout=layout)
    e2 = ctx.event()
    tf = ctx.elapsed_ms(e0, e1, destroy=False)
    tb = ctx.elapsed_ms(e1, e2, destroy=False)
    print(f"rep {rep}: filter {tf:.3f} ms ({8.0*a.det*a.det*n/tf/1e6:.1f} GB/s)  backproject {tb:.3f} ms ({updates/tb/1e6:.1f} GUPS)")
if a.check:
    out = np.empty((vol.dim_z, vol.dim_y, vol.dim_x), np.float32)
    ctx.vol_d2h(d_vol, out, out.size)
    ctx.set_option("bp_kernel", 1)
    d_ref = ctx.volume_alloc(*dims)
    for rep in range(a.reps):
        ctx.backproject_stack(stack, 0, n, sc[:, 0], sc[:, 1], d_ref, dims, 0, det, vol, layout=layout)
    ref = np.empty_like(out)
    ctx.vol_d2h(d_ref, ref, ref.size)
    c = n / (8 * np.pi) * a.reps
    d = out.astype(np.float64) - ref
    print(f"vs exact kernel: max {np.abs(d).max()/c:.3e} C, rmse {np.sqrt((d*d).mean())/c:.3e} C; launches {ctx.launch_count()}")
