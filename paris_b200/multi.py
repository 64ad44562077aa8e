"""z-slab reconstruction across the GPUs of one box: one process per GPU, NCCL all-gather of the
filtered stack (the slab arithmetic is the reference's, src/cuda/subvolume_information.cpp:112-116,
src/make_volume.cpp:32-34, src/main.cpp:96).

    rank r   uploads and filters projections [r*chunk, (r+1)*chunk)   (fused weight+filter kernel)
    all      all-gather of the filtered stack over NVLink (torch.distributed / NCCL), in place
    rank r   backprojects ALL projections into z-slab r -- no reduction; the host reassembles slabs by offset

torch is plumbing here (device memory for the stack that NCCL can see, the process group); every
kernel is launched by libparis_b200.so on its own stream.  With world == 1 this degenerates to the
single-GPU pipeline and the end-to-end step runs through the C++ per-projection loop.
"""
from __future__ import annotations

import numpy as np

from . import capi, dropin
from .pipeline import angle_sin_cos


class SlabPlan:
    """Equal z-slabs, remainder on the last one; contiguous, equal blocks of projections per rank."""

    def __init__(self, dim_z: int, world: int, rank: int):
        self.world, self.rank = world, rank
        self.dz = dim_z // world
        self.remainder = dim_z % world
        self.offset = rank * self.dz                                   # src/main.cpp:96
        self.slab_dz = self.dz + (self.remainder if rank == world - 1 else 0)  # src/make_volume.cpp:32-34

    def projection_block(self, n_proj: int):
        """[lo, hi) of the projections this rank uploads and filters, and the common chunk length."""
        chunk = (n_proj + self.world - 1) // self.world
        lo = min(self.rank * chunk, n_proj)
        hi = min(lo + chunk, n_proj)
        return lo, hi, chunk


class MultiGpuReconstructor:
    def __init__(self, device: int, det: capi.DetectorGeometry, vol: capi.VolumeGeometry, n_proj: int,
                 plan: SlabPlan, dist=None, batch: int = 64):
        self.det, self.vol, self.n_proj, self.plan, self.dist = det, vol, n_proj, plan, dist
        self.device = device
        # the C++ layer's per-thread context, so the e2e loop and the stack-level calls share streams
        dropin.set_device(device)
        self.ctx = capi.Context(device, handle=dropin.context_handle())
        self.batch = batch
        self.ctx.set_option("bp_batch", batch)
        self.px = det.n_row * det.n_col
        self.lo, self.hi, self.chunk = plan.projection_block(n_proj)
        self.my_count = self.hi - self.lo
        self.slot_bytes, self.pitch = capi.stack_slot_bytes(det.n_row, det.n_col)
        self.layout = capi.choose_stack_layout(det, vol)
        self.slots = self.chunk * plan.world
        self.filter = self.ctx.filter_create(capi.filter_size(det.n_row), float(det.l_px_row))
        sc = np.array([angle_sin_cos(i, det) for i in range(n_proj)], dtype=np.float32).reshape(n_proj, 2)
        self.sin = np.ascontiguousarray(sc[:, 0])
        self.cos = np.ascontiguousarray(sc[:, 1])
        self.slab_dims = (vol.dim_x, vol.dim_y, plan.slab_dz)
        self.d_vol = self.ctx.volume_alloc(*self.slab_dims)
        self._torch_stack = None
        if dist is not None:
            import torch
            self._torch = torch
            self._torch_stack = torch.zeros(self.slots * self.slot_bytes // 4, dtype=torch.float32,
                                            device=torch.device("cuda", device))
            self.d_stack = self._torch_stack.data_ptr()
            self._ext_stream = torch.cuda.ExternalStream(self.ctx.stream(), device=torch.device("cuda", device))
        else:
            self.d_stack = self.ctx.dev_alloc(self.slots * self.slot_bytes)
        self.d_raw = None
        self.h_raw = None
        self.h_slab = capi.PinnedArray((plan.slab_dz, vol.dim_y, vol.dim_x))

    # ---- inputs -----------------------------------------------------------------------------------------------
    def generate_inputs(self, ellipsoids_mm: np.ndarray):
        """Synthetic raw projections [lo, hi) on the device (phantom kernel) and mirrored in pinned host memory."""
        n = max(self.my_count, 1)
        self.d_raw = self.ctx.dev_alloc(n * self.px * 4)
        self.h_raw = capi.PinnedArray((n, self.det.n_col, self.det.n_row))
        if self.my_count:
            self.ctx.phantom_project(ellipsoids_mm, self.det, self.lo, self.my_count, self.d_raw)
            for i in range(self.my_count):
                self.ctx.proj_d2h(self.d_raw + i * self.px * 4, self.h_raw.ptr + i * self.px * 4, self.det.n_row,
                                  self.det.n_col)
        self.ctx.sync()

    def host_sample(self, count: int, stride: int = 1) -> np.ndarray:
        return np.ascontiguousarray(self.h_raw.array[::stride][:count])

    # ---- exchange ----------------------------------------------------------------------------------------------
    def _allgather(self):
        if self.dist is None:
            return
        torch = self._torch
        chunk_floats = self.chunk * self.slot_bytes // 4
        mine = self._torch_stack[self.plan.rank * chunk_floats:(self.plan.rank + 1) * chunk_floats]
        with torch.cuda.stream(self._ext_stream):
            self.dist.all_gather_into_tensor(self._torch_stack, mine)

    # ---- steps -------------------------------------------------------------------------------------------------
    def step_resident(self, timed: bool = False):
        """raw projections already in HBM -> slab in HBM.  Returns (filter, all-gather, backproject) ms if timed."""
        ctx = self.ctx
        e0 = ctx.event() if timed else None
        ctx.volume_clear(self.d_vol, *self.slab_dims)
        if self.my_count:
            ctx.filter_to_stack_batch(self.d_raw, self.px, self.my_count, self.det, self.filter, self.d_stack, self.lo,
                                      self.layout)
        e1 = ctx.event() if timed else None
        self._allgather()
        e2 = ctx.event() if timed else None
        ctx.backproject_stack(self.d_stack, 0, self.n_proj, self.sin, self.cos, self.d_vol, self.slab_dims,
                              self.plan.offset, self.det, self.vol, layout=self.layout)
        if not timed:
            return None
        e3 = ctx.event()
        tf = ctx.elapsed_ms(e0, e1, destroy=False)
        tg = ctx.elapsed_ms(e1, e2, destroy=False)
        tb = ctx.elapsed_ms(e2, e3, destroy=False)
        for e in (e0, e1, e2, e3):
            capi.check(capi.lib().paris_b200_event_destroy(e))
        return tf, tg, tb

    def step_e2e(self):
        """pinned host raw projections -> pinned host slab, copies included."""
        if self.dist is None:
            # the reference-shaped per-projection loop in C++ (paris_b200/cpp/pipeline.cpp: reconstruct_task)
            dropin.reconstruct(self.h_raw.ptr, self.n_proj, self.det, self.vol, self.h_slab.ptr,
                               (self.vol.dim_x, self.vol.dim_y, self.vol.dim_z), device=self.device)
            return
        ctx = self.ctx
        ctx.volume_clear(self.d_vol, *self.slab_dims)
        for i in range(self.my_count):
            d = ctx.dev_alloc(self.px * 4)
            ctx.proj_h2d(self.h_raw.ptr + i * self.px * 4, d, self.det.n_row, self.det.n_col)
            ctx.filter_to_stack(d, self.det, self.filter, self.d_stack, self.lo + i, self.layout)
            ctx.dev_free(d)
        self._allgather()
        ctx.backproject_stack(self.d_stack, 0, self.n_proj, self.sin, self.cos, self.d_vol, self.slab_dims,
                              self.plan.offset, self.det, self.vol, layout=self.layout)
        ctx.vol_d2h(self.d_vol, self.h_slab.ptr, self.slab_dims[0] * self.slab_dims[1] * self.slab_dims[2])

    def slab(self) -> np.ndarray:
        return self.h_slab.array

    def close(self):
        self.ctx.sync()
        self.ctx.filter_destroy(self.filter)
        self.ctx.volume_free(self.d_vol)
        if self.dist is None:
            self.ctx.dev_free(self.d_stack)
        if self.d_raw is not None:
            self.ctx.dev_free(self.d_raw)
        if self.h_raw is not None:
            self.h_raw.free()
        self.h_slab.free()
        self._torch_stack = None
