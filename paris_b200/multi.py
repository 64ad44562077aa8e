"""One scan across the GPUs of one box: Python plumbing over the group entry points of the C ABI
(include/paris_b200.h, csrc/group.cu).

Everything that matters happens in C++/CUDA behind ``paris_b200_group_*``: member r uploads and filters 1/world of
the projections round by round, copies the detector-row band each peer needs straight into that peer's stack over
NVLink (peer memory, copy engines, arrival flags the peer's backprojection stream waits on), backprojects all
projections into its own z-slabs and streams them to the host.  This module only

  * generates a member's synthetic raw projections (phantom kernel) and mirrors them in pinned host memory,
  * hands the members' handles around (the caller brings the channel: torch.distributed in bench.py, a pipe in the
    tests, nothing at all for members of one process),
  * offers the slab arithmetic of the reference (src/cuda/subvolume_information.cpp:112-116, src/main.cpp:96,
    src/make_volume.cpp:32-34) to callers that plan on the host.
"""
from __future__ import annotations

import numpy as np

from . import capi


class SlabPlan:
    """Equal z-slabs, remainder on the last one (what group_plan computes, for one slab per member)."""

    def __init__(self, dim_z: int, world: int, rank: int):
        if world > dim_z:
            raise ValueError(f"{world} ranks cannot share {dim_z} slices: every rank needs at least one "
                             "(run with at most as many ranks as the region has slices)")
        self.world, self.rank = world, rank
        self.dz = dim_z // world
        self.remainder = dim_z % world
        self.offset = rank * self.dz                                   # src/main.cpp:96
        self.slab_dz = self.dz + (self.remainder if rank == world - 1 else 0)  # src/make_volume.cpp:32-34


class GroupMember:
    """A group member with its synthetic inputs: what bench.py and the tests drive."""

    def __init__(self, device: int, rank: int, world: int, det: capi.DetectorGeometry, vol: capi.VolumeGeometry,
                 n_proj: int, roi: capi.Roi | None = None, **options):
        """vol: the FULL volume geometry; roi: the reconstructed box (default: the whole volume); options: the
        remaining fields of capi.group_config (slabs_per_member, stream_slabs, first_round, max_round,
        whole_projections, exchange, angles_deg, x_parts, host_row_floats, sample_type)."""
        self.det, self.vol, self.n_proj, self.rank, self.world, self.device = det, vol, n_proj, rank, world, device
        self.cfg = capi.group_config(rank, world, det, vol, n_proj, roi=roi, **options)
        self.plan = capi.group_plan(self.cfg)
        self.group = capi.Group(device, self.cfg)
        self.info = self.group.info()
        self.ctx = capi.Context(device, handle=self.info.ctx)          # the backprojection context (events, options)
        self.fctx = capi.Context(device, handle=self.info.filter_ctx)
        self.px = det.n_row * det.n_col
        self.sample_bytes = 2 if self.cfg.sample_type == capi.SAMPLES_U16 else 4
        self.proj_bytes = self.px * self.sample_bytes
        self.my_count = self.info.my_projections
        self.slice_floats = self.info.x_count * self.info.region_y      # of the member's device slabs
        # this member's projections: per round one run of consecutive scan indices
        self.runs = []
        for rd in range(self.plan.rounds):
            first, count = capi.group_share(self.plan, world, rd, rank)
            if count:
                self.runs.append((first, count))
        self.d_raw = None
        self.h_raw = None
        self.h_slabs = None

    # ---- handles ---------------------------------------------------------------------------------------------------
    def export(self) -> bytes:
        return self.group.export()

    def connect(self, handles):
        self.group.connect(list(handles))

    # ---- inputs ----------------------------------------------------------------------------------------------------
    def projection_indices(self):
        return [i for first, count in self.runs for i in range(first, first + count)]

    def generate_inputs(self, ellipsoids_mm: np.ndarray, host: bool = True, counts_scale: float | None = None):
        """This member's raw projections on the device (phantom kernel, local order) and, for the end-to-end steps,
        mirrored in pinned host memory.  counts_scale: turn the line integrals into detector counts,
        rint(value * scale) clipped to 16 bits -- what a 16-bit member needs, and what a float member is given when
        the two are compared (the same numbers in the other sample type)."""
        n = max(self.my_count, 1)
        d_float = self.ctx.dev_alloc(n * self.px * 4)
        local = 0
        for first, count in self.runs:
            self.ctx.phantom_project(ellipsoids_mm, self.det, first, count, d_float + local * self.px * 4)
            local += count
        if counts_scale is None:
            if self.sample_bytes != 4:
                raise ValueError("16-bit samples need counts_scale")
            self.d_raw = d_float
            if host:
                self.h_raw = capi.PinnedArray((n, self.det.n_col, self.det.n_row))
                for i in range(self.my_count):
                    self.ctx.proj_d2h(d_float + i * self.px * 4, self.h_raw.ptr + i * self.px * 4, self.det.n_row,
                                      self.det.n_col)
            self.ctx.sync()
            return
        dtype = np.uint16 if self.sample_bytes == 2 else np.float32
        self.h_raw = capi.PinnedArray((n, self.det.n_col, self.det.n_row), dtype=dtype)
        line = np.empty((self.det.n_col, self.det.n_row), np.float32)
        for i in range(self.my_count):
            self.ctx.proj_d2h(d_float + i * self.px * 4, line, self.det.n_row, self.det.n_col)
            self.ctx.sync()
            self.h_raw.array[i] = np.clip(np.rint(line * np.float32(counts_scale)), 0, 65535).astype(dtype)
        self.ctx.dev_free(d_float)
        self.d_raw = self.ctx.dev_alloc(n * self.proj_bytes)
        self.ctx.vol_h2d(self.h_raw.ptr, self.d_raw, n * self.proj_bytes // 4)
        self.ctx.sync()

    def host_sample(self, count: int, stride: int = 1) -> np.ndarray:
        """(count, n_col, n_row) raw projections, every stride-th of this member's (world == 1: of the scan)"""
        return np.ascontiguousarray(self.h_raw.array[::stride][:count], dtype=np.float32)

    def host_pointers(self):
        return [self.h_raw.ptr + i * self.proj_bytes for i in range(self.my_count)]

    def alloc_host_slabs(self):
        """pinned host memory of this member's own (for callers without a shared host volume)"""
        self.h_slabs = capi.PinnedArray((self.info.z_count, self.info.region_y, self.info.x_count))
        return self.h_slabs

    # ---- steps -----------------------------------------------------------------------------------------------------
    def begin_resident(self):
        self.group.begin(d_raw=self.d_raw)

    def begin_e2e(self, h_slabs_ptr: int | None = None):
        if h_slabs_ptr is None:
            h_slabs_ptr = self.h_slabs.ptr
        self.group.begin(h_raw=self.host_pointers(), h_slabs=h_slabs_ptr)

    def end(self):
        self.group.end()

    def step_resident(self):
        self.begin_resident()
        self.end()

    def step_e2e(self, h_slabs_ptr: int | None = None):
        self.begin_e2e(h_slabs_ptr)
        self.end()

    def launch_count(self) -> int:
        return self.ctx.launch_count() + self.fctx.launch_count()

    def device_slab(self, z_first: int, dz: int) -> np.ndarray:
        """slices [z_first, z_first + dz) of this member's FIRST slab, from the device (resident slabs)"""
        out = np.empty((dz, self.info.region_y, self.info.x_count), np.float32)
        self.ctx.vol_d2h(self.info.d_first_slab + z_first * self.slice_floats * 4, out, out.size)
        return out

    def box_roi(self, roi: capi.Roi | None) -> capi.Roi | None:
        """The ROI that shifts voxel indices to this member's columns (stack-level calls outside the group)."""
        if self.info.x_first == 0:
            return roi
        r = capi.Roi(0, 0, 0, 0, 0, 0) if roi is None else capi.Roi(roi.x1, roi.x2, roi.y1, roi.y2, roi.z1, roi.z2)
        r.x1 += self.info.x_first
        return r

    def host_offset_bytes(self, region_x: int) -> int:
        """Where this member's box starts inside a region-wide host volume of `region_x` floats per row."""
        return (self.info.z_first * self.info.region_y * region_x + self.info.x_first) * 4

    def close(self):
        self.group.end()
        if self.d_raw is not None:
            self.ctx.dev_free(self.d_raw)
            self.d_raw = None
        if self.h_raw is not None:
            self.h_raw.free()
            self.h_raw = None
        if self.h_slabs is not None:
            self.h_slabs.free()
            self.h_slabs = None
        self.group.close()
