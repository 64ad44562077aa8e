"""Developer probe: host<->device copy bandwidth through the library's own copy paths."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from paris_b200 import capi
ctx = capi.Context(0)
n, px = 256, 1024 * 1024
h = capi.PinnedArray((n, 1024, 1024))
h.array[:] = 1.0
bufs = [ctx.dev_alloc(px * 4) for _ in range(n)]
for rep in range(3):
    ctx.sync()
    t0 = time.perf_counter()
    for i in range(n):
        ctx.proj_h2d(h.ptr + i * px * 4, bufs[i], 1024, 1024)
    ctx.sync()
    dt = time.perf_counter() - t0
    print(f"H2D pinned, {n} x 4 MiB: {n*px*4/dt/1e9:.1f} GB/s")
vol = ctx.volume_alloc(512, 512, 512)
hv = capi.PinnedArray((512, 512, 512))
for rep in range(3):
    t0 = time.perf_counter()
    ctx.vol_d2h(vol, hv.ptr, 512 ** 3)
    dt = time.perf_counter() - t0
    print(f"D2H pinned 512 MiB: {512**3*4/dt/1e9:.1f} GB/s")
pg = np.ones((64, 1024, 1024), np.float32)
t0 = time.perf_counter()
for i in range(64):
    ctx.proj_h2d(pg[i], bufs[i], 1024, 1024)
ctx.sync()
dt = time.perf_counter() - t0
print(f"H2D pageable: {64*px*4/dt/1e9:.1f} GB/s")
os.system("nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max,pcie.link.width.max --format=csv")
os.system("nproc; lscpu | grep -E 'Model name|Socket|NUMA node\\(s\\)' ")
