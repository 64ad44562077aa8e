"""Small cases for a checked run.  compute-sanitizer is CLOSED on the GPU pool (gpurun refuses it: "runs under it have
left GPUs needing a reset"), so the check is the library's own: `make check-lib` builds
paris_b200/libparis_b200_check.so, whose backprojection tests every shared-memory sample load against the CTA's ring of
staged boxes; run this script with PARIS_B200_LIB pointing at it and it prints the counters.  The cases (also what a
compute-sanitizer run would use, ONE tool per gpurun call, where that tool is open):
  a  config-1-like coarse volume: fused weight+filter, parity-split stack, TMA kernel (interior, MIXED, skipped tiles)
  b  natural volume, ROI with odd offsets, detector shifted, three z-slabs: plain stack, STRADDLE tiles, chunked download
  c  a reconstruction group of two members sharing the GPU (bands pushed by the copy engine and by the copy kernel,
     flag kernels)
Prints the max error of each against the one-piece / exact results so that a sanitizer run is also a numerics run."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from paris_b200 import capi
import test_gpu_group as T

which = sys.argv[1] if len(sys.argv) > 1 else "abc"
ctx = capi.Context(0)
if "a" in which:
    det, vol, n_proj = T._case(k=48, extra_z=0, n=96, n_proj=24)
    fast = T._one_piece(ctx, det, vol, n_proj)
    print("a: coarse 48^3 from 24 x 96^2, max |v|", float(np.abs(fast).max()), ctx.bp_kernel_info()["last"], flush=True)
if "b" in which:
    n, n_proj = 96, 20
    det = capi.DetectorGeometry(n, 80, 0.4, 0.4, 4.0, 0.0, 500.0, 500.0, 360.0 / n_proj)
    vol = capi.calculate_volume_geometry(det)
    roi = capi.Roi(9, 71, 14, 88, 5, 5 + 61)
    reg = capi.apply_roi(vol, roi)
    region = (reg.dim_x, reg.dim_y, reg.dim_z)
    want = T._one_piece(ctx, det, vol, n_proj, roi=roi, region=region)
    got, _ = T._run_in_process(1, det, vol, n_proj, roi=roi, slabs_per_member=3, stream_slabs=True, first_round=5, max_round=10)
    print("b: ROI", region, "three streamed slabs equal one piece:", bool(np.array_equal(got, want)), ctx.bp_kernel_info()["last"], flush=True)
if "c" in which:
    det, vol, n_proj = T._case(k=32, extra_z=1, n=64, n_proj=16)
    want = T._one_piece(ctx, det, vol, n_proj)
    for ex in (capi.EXCHANGE_COPY_ENGINE, capi.EXCHANGE_KERNEL):
        got, stats = T._run_in_process(2, det, vol, n_proj, first_round=4, max_round=8, exchange=ex)
        print("c: two members, exchange", ex, "equal one piece:", bool(np.array_equal(got, want)), flush=True)
if "d" in which:
    # config 1 at full size (128^3 from 256 x 256^2): tall split tiles, every kind of tile
    det, vol, n_proj = T._case(k=128, extra_z=0, n=256, n_proj=256)
    fast = T._one_piece(ctx, det, vol, n_proj)
    print("d: config 1, 128^3 from 256 x 256^2, max |v|", float(np.abs(fast).max()), ctx.bp_kernel_info()["last"], flush=True)
import ctypes
L = capi.lib()
if hasattr(L, "paris_b200_debug_bounds"):
    out = (ctypes.c_ulonglong * 4)()
    L.paris_b200_debug_bounds(out)
    print(f"bounds check: {out[0]} sample loads checked, {out[1]} outside the staged ring"
          + (f" (offsets {ctypes.c_longlong(out[2]).value} .. {out[3]})" if out[1] else ""), flush=True)
    assert out[0] > 0 and out[1] == 0
else:
    print("(unchecked library: build `make check-lib` and set PARIS_B200_LIB)", flush=True)
ctx.close()
print("done", flush=True)
