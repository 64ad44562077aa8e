"""Where does the N=8 step of config 3 lose its 8 %?  One GPU plays every member in turn: the whole stack is filtered
once, then all 1440 projections are backprojected into each of the eight 128-slice slabs on its own (what a member's
backprojection stream does, without any exchange or filter next to it) and into the whole volume.

    python scripts/r2_slab_times.py [--config c3] [--parts 8] > gpurun_out/r2_slab_times.json
"""
import argparse, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from paris_b200 import capi
from paris_b200.multi import GroupMember
from paris_b200.pipeline import angle_sin_cos

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c3")
ap.add_argument("--parts", type=int, default=8)
ap.add_argument("--reps", type=int, default=2)
args = ap.parse_args()
det, vol, n_proj, roi, dims = bench.geometry(args.config)
m = GroupMember(0, 0, 1, det, vol, n_proj, roi=roi)
m.generate_inputs(bench.ellipsoids(det), host=False)
m.step_resident()                        # fills the stack
ctx, info = m.ctx, m.info
sc = np.array([angle_sin_cos(i, det) for i in range(n_proj)], dtype=np.float32)


def timed(slab_dims, z_first, count=n_proj):
    best = 1e30
    for _ in range(args.reps + 1):
        ctx.volume_clear(info.d_first_slab, *slab_dims)
        e0 = ctx.event()
        ctx.backproject_stack(info.d_stack, 0, count, sc[:, 0], sc[:, 1], info.d_first_slab, slab_dims, z_first, det, vol,
                              roi=roi, layout=info.layout)
        e1 = ctx.event()
        best = min(best, ctx.elapsed_ms(e0, e1))
    return best


whole = timed(dims, 0)
dz = dims[2] // args.parts
slabs = [timed((dims[0], dims[1], dz), k * dz) for k in range(args.parts)]
one_batch = [timed((dims[0], dims[1], dz), k * dz, count=256) for k in (0, args.parts // 2)]
small_batch = [timed((dims[0], dims[1], dz), k * dz, count=64) for k in (0, args.parts // 2)]
print(json.dumps({"config": args.config, "whole_volume_ms": whole, "slab_slices": dz, "slab_ms": slabs,
                  "sum_of_slabs_over_whole": sum(slabs) / whole, "slowest_slab_times_parts_over_whole": max(slabs) * args.parts / whole,
                  "one_256_batch_ms_slab0_mid": one_batch, "one_64_batch_ms_slab0_mid": small_batch,
                  "kernel": ctx.bp_kernel_info()["last"]}))
m.close()
