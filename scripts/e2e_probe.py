"""Developer probe: where does the end-to-end (host -> host) time go?"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from paris_b200 import capi, phantom, dropin
from paris_b200.pipeline import Pipeline, angle_sin_cos
import bench

det, vol, n_proj = bench.geometry("c2")
dropin.set_device(0)
ctx = capi.Context(0, handle=dropin.context_handle())
px = det.n_row * det.n_col
d_raw = ctx.dev_alloc(n_proj * px * 4)
ctx.phantom_project(bench.ellipsoids(det), det, 0, n_proj, d_raw)
h = capi.PinnedArray((n_proj, det.n_col, det.n_row))
for i in range(n_proj):
    ctx.proj_d2h(d_raw + i * px * 4, h.ptr + i * px * 4, det.n_row, det.n_col)
hv = capi.PinnedArray((vol.dim_z, vol.dim_y, vol.dim_x))
dims = (vol.dim_x, vol.dim_y, vol.dim_z)

def t(label, fn, reps=3):
    for r in range(reps):
        ctx.sync(); t0 = time.perf_counter(); fn(); ctx.sync(); dt = time.perf_counter() - t0
        print(f"{label:50s} rep {r} {dt*1e3:8.1f} ms  {ctx.stats()}")

def h2d_only():
    for i in range(n_proj):
        d = ctx.dev_alloc(px * 4); ctx.proj_h2d(h.ptr + i * px * 4, d, det.n_row, det.n_col); ctx.dev_free(d)
t("python loop: alloc + h2d + free", h2d_only)

pl = Pipeline(ctx, det)
v = pl.make_volume(*dims)
def full_py():
    ctx.volume_clear(v.d_ptr, *dims)
    for i in range(n_proj):
        d = ctx.dev_alloc(px * 4); ctx.proj_h2d(h.ptr + i * px * 4, d, det.n_row, det.n_col)
        sn, cs = sc[i]
        ctx.backproject(d, det.n_row, det.n_col, v.d_ptr, dims, 0, det, vol, None, sn, cs, 0.0, 0.0, capi.BP_FUSE_WEIGHT_FILTER, pl.filter_handle)
        ctx.dev_free(d)
    ctx.flush()
sc = [angle_sin_cos(i, det) for i in range(n_proj)]
t("python loop: + fused backproject (deferred)", full_py)
def full_py_d2h():
    full_py(); ctx.vol_d2h(v.d_ptr, hv.ptr, hv.array.size)
t("python loop: + vol d2h", full_py_d2h)
t("C++ loop (dropin.reconstruct)", lambda: dropin.reconstruct(h.ptr, n_proj, det, vol, hv.ptr, dims), reps=5)
ctx.set_option("bp_batch", 32)
t("C++ loop, bp_batch 32", lambda: dropin.reconstruct(h.ptr, n_proj, det, vol, hv.ptr, dims))
ctx.set_option("bp_batch", 16)
t("C++ loop, bp_batch 16", lambda: dropin.reconstruct(h.ptr, n_proj, det, vol, hv.ptr, dims))
