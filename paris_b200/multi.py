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

import os
import time

import numpy as np

from . import capi, dropin
from .pipeline import angle_sin_cos


class SlabPlan:
    """Equal z-slabs, remainder on the last one; contiguous, equal blocks of projections per rank."""

    def __init__(self, dim_z: int, world: int, rank: int):
        if world > dim_z:
            raise ValueError(f"{world} ranks cannot share {dim_z} slices: every rank needs at least one "
                             "(run with at most as many ranks as the region has slices)")
        self.world, self.rank = world, rank
        self.dz = dim_z // world
        self.remainder = dim_z % world
        self.offset = rank * self.dz                                   # src/main.cpp:96
        self.slab_dz = self.dz + (self.remainder if rank == world - 1 else 0)  # src/make_volume.cpp:32-34

    def projection_block(self, n_proj: int):
        """[lo, hi) of the projections this rank uploads and filters, and the common chunk length
        (contiguous assignment: one all-gather of the whole stack)."""
        chunk = (n_proj + self.world - 1) // self.world
        lo = min(self.rank * chunk, n_proj)
        hi = min(lo + chunk, n_proj)
        return lo, hi, chunk

    def cyclic_blocks(self, n_proj: int, max_round: int = 64):
        """Block-cyclic assignment for the pipelined exchange: the scan is cut into rounds of world*m
        consecutive projections, rank r owns the m projections [round*world*m + r*m, ... + m) of every round,
        so a round's all-gather output is a contiguous run of stack slots in PROJECTION ORDER (the
        backprojection then adds projections in the same order as a single GPU does).  Returns m (0 if the
        scan does not divide evenly: callers fall back to the contiguous scheme)."""
        if n_proj % self.world:
            return 0
        per_rank = n_proj // self.world
        best = 0
        for m in range(1, per_rank + 1):
            if per_rank % m == 0 and self.world * m <= max_round:
                best = m
        return best


def round_schedule(n_proj: int, world: int, first_round: int = 64, max_round: int = 256):
    """Projections per rank and round for the pipelined exchange: rounds of world*m consecutive projections with m
    growing 2x from first_round/world up to max_round/world -- short rounds first so that the backprojection starts
    early, long ones afterwards (few large all-gathers overlap with the backprojection far better than many small
    ones).  Empty if the scan does not divide evenly among the ranks (callers fall back to one big all-gather)."""
    if n_proj % world:
        return []
    remaining, cap, ms = n_proj // world, max(first_round, world), []
    while remaining > 0:
        m = min(remaining, max(1, cap // world))
        if remaining - m < max(1, m // 4):   # do not leave a sliver for the last round
            m = remaining
        ms.append(m)
        remaining -= m
        cap = min(max(max_round, world), cap * 2)
    return ms


class MultiGpuReconstructor:
    def __init__(self, device: int, det: capi.DetectorGeometry, vol: capi.VolumeGeometry, n_proj: int,
                 plan: SlabPlan, dist=None, batch: int = 256, roi: capi.Roi | None = None, region=None,
                 gather_round: int = 64):
        """vol: the FULL volume geometry; roi/region: the reconstructed box and its (x, y, z) dimensions (default:
        the whole volume).  plan cuts the REGION's z extent."""
        self.det, self.vol, self.n_proj, self.plan, self.dist = det, vol, n_proj, plan, dist
        self.roi = roi
        self.region = tuple(region) if region is not None else (vol.dim_x, vol.dim_y, vol.dim_z)
        self.device = device
        # the C++ layer's per-thread context, so the e2e loop and the stack-level calls share streams
        dropin.set_device(device)
        self.ctx = capi.Context(device, handle=dropin.context_handle())
        batch = int(os.environ.get("PARIS_B200_BATCH", batch))   # (tuning experiments)
        self.batch = batch
        self.ctx.set_option("bp_batch", batch)
        self.px = det.n_row * det.n_col
        self.lo, self.hi, self.chunk = plan.projection_block(n_proj)
        self.my_count = self.hi - self.lo
        # pipelined exchange (N > 1): block-cyclic ownership, m projections per rank and round
        # (the first exchanged round holds at most `gather_round` projections, later ones up to `batch`)
        gather_round = int(os.environ.get("PARIS_B200_GATHER_ROUND", gather_round))
        self.ms = round_schedule(n_proj, plan.world, min(batch, gather_round), batch) if dist is not None else []
        self.rounds = len(self.ms)
        self.m = self.ms[0] if self.ms else 0                       # (non-zero = pipelined exchange available)
        self.local_start = [sum(self.ms[:r]) for r in range(self.rounds + 1)]   # first local projection of round r
        self.slot_bytes, self.pitch = capi.stack_slot_bytes(det.n_row, det.n_col)
        self.layout = capi.choose_stack_layout(det, vol)
        self.slots = self.chunk * plan.world
        self.filter = self.ctx.filter_create(capi.filter_size(det.n_row), float(det.l_px_row))
        sc = np.array([angle_sin_cos(i, det) for i in range(n_proj)], dtype=np.float32).reshape(n_proj, 2)
        self.sin = np.ascontiguousarray(sc[:, 0])
        self.cos = np.ascontiguousarray(sc[:, 1])
        self.slab_dims = (self.region[0], self.region[1], plan.slab_dz)
        self.d_vol = self.ctx.volume_alloc(*self.slab_dims)
        self._torch_stack = None
        if dist is not None:
            import torch
            self._torch = torch
            self._torch_stack = torch.zeros(self.slots * self.slot_bytes // 4, dtype=torch.float32,
                                            device=torch.device("cuda", device))
            self.d_stack = self._torch_stack.data_ptr()
            self._ext_stream = torch.cuda.ExternalStream(self.ctx.stream(), device=torch.device("cuda", device))
            self._comm_stream = torch.cuda.Stream(device=torch.device("cuda", device), priority=-1)
            # uploads and the fused weight+filter kernel run in a context of their own (own streams, own buffer
            # pool): a stream is in-order, so filter launches queued behind a long backprojection would hold back
            # the exchange the NEXT backprojection waits for, and a backprojection queued behind a filter launch
            # would wait for that filter's upload
            self.fctx = capi.Context(device)
            self.fctx.set_option("bp_batch", batch)   # (sizes its projection buffer pool)
            self._filter_stream = torch.cuda.ExternalStream(self.fctx.stream(), device=torch.device("cuda", device))
        else:
            self.d_stack = self.ctx.stack_alloc(det.n_row, det.n_col, self.slots)
            self.fctx = self.ctx
        self.d_raw = None
        self.h_raw = None
        self.h_slab = capi.PinnedArray((plan.slab_dz, self.region[1], self.region[0]))

    # ---- inputs -----------------------------------------------------------------------------------------------
    def generate_inputs(self, ellipsoids_mm: np.ndarray):
        """Synthetic raw projections [lo, hi) on the device (phantom kernel) and mirrored in pinned host memory."""
        n = max(self.my_count, 1)
        self.d_raw = self.ctx.dev_alloc(n * self.px * 4)
        self.h_raw = capi.PinnedArray((n, self.det.n_col, self.det.n_row))
        if self.my_count:
            if self.m:
                # local projection i = round*m + j  <->  global index round*world*m + rank*m + j
                for rd in range(self.rounds):
                    self.ctx.phantom_project(ellipsoids_mm, self.det, self.global_index(self.local_start[rd]),
                                             self.ms[rd], self.d_raw + self.local_start[rd] * self.px * 4)
            else:
                self.ctx.phantom_project(ellipsoids_mm, self.det, self.lo, self.my_count, self.d_raw)
            for i in range(self.my_count):
                self.ctx.proj_d2h(self.d_raw + i * self.px * 4, self.h_raw.ptr + i * self.px * 4, self.det.n_row,
                                  self.det.n_col)
        self.ctx.sync()

    def launch_count(self) -> int:
        """Kernel launches issued so far by this reconstructor's context(s)."""
        n = self.ctx.launch_count()
        if self.fctx is not self.ctx:
            n += self.fctx.launch_count()
        return n

    def global_index(self, local: int) -> int:
        """Scan index (= stack slot) of this rank's local projection `local`."""
        if not self.m:
            return self.lo + local
        rd = max(r for r in range(self.rounds) if self.local_start[r] <= local)
        j = local - self.local_start[rd]
        return self.plan.world * self.local_start[rd] + self.plan.rank * self.ms[rd] + j

    def host_sample(self, count: int, stride: int = 1) -> np.ndarray:
        return np.ascontiguousarray(self.h_raw.array[::stride][:count])

    # ---- exchange ----------------------------------------------------------------------------------------------
    def _allgather(self):
        if self.dist is None:
            return
        torch = self._torch
        if self.m:
            slot_floats = self.slot_bytes // 4
            w = self.plan.world
            with torch.cuda.stream(self._ext_stream):
                for rd in range(self.rounds):
                    m = self.ms[rd]
                    first = w * self.local_start[rd]
                    mine_first = first + self.plan.rank * m
                    self.dist.all_gather_into_tensor(self._torch_stack[first * slot_floats:(first + w * m) * slot_floats],
                                                     self._torch_stack[mine_first * slot_floats:(mine_first + m) * slot_floats])
            return
        chunk_floats = self.chunk * self.slot_bytes // 4
        mine = self._torch_stack[self.plan.rank * chunk_floats:(self.plan.rank + 1) * chunk_floats]
        with torch.cuda.stream(self._ext_stream):
            self.dist.all_gather_into_tensor(self._torch_stack, mine)

    # ---- steps -------------------------------------------------------------------------------------------------
    def _backproject(self, first: int, count: int, download: bool = False):
        if download:
            # last round of an end-to-end step: the slab goes to the host chunk by chunk behind the kernel
            self.ctx.backproject_stack_d2h(self.d_stack, first, count, self.sin[first:first + count],
                                           self.cos[first:first + count], self.d_vol, self.slab_dims, self.plan.offset,
                                           self.det, self.vol, self.h_slab.ptr, roi=self.roi, layout=self.layout)
            return
        self.ctx.backproject_stack(self.d_stack, first, count, self.sin[first:first + count], self.cos[first:first + count],
                                   self.d_vol, self.slab_dims, self.plan.offset, self.det, self.vol, roi=self.roi,
                                   layout=self.layout)

    def _pipelined(self, upload: bool):
        """N > 1: per round, upload + filter my m projections (filter context) -> all-gather the round (comm stream)
        -> backproject groups of rounds (compute context).  Three independent streams ordered only by events, so the
        exchange of a round hides behind the backprojection of the rounds before it and nothing waits for an upload
        it does not need.  Rounds grow from `gather_round` to one batch of projections (round_schedule)."""
        ctx, fctx, torch = self.ctx, self.fctx, self._torch
        w = self.plan.world
        slot_floats = self.slot_bytes // 4
        trace = os.environ.get("PARIS_B200_TRACE") and upload
        t_begin = time.perf_counter()
        # the stack slots are rewritten: the previous step's backprojection must have read them
        prev_done = torch.cuda.Event()
        prev_done.record(self._ext_stream)
        self._filter_stream.wait_event(prev_done)
        ctx.volume_clear(self.d_vol, *self.slab_dims)
        ev_trace = os.environ.get("PARIS_B200_TRACE_EVENTS")
        if ev_trace:
            t_ev0 = torch.cuda.Event(enable_timing=True)
            t_ev0.record(self._ext_stream)
            comm_marks = []
        for rd in range(self.rounds):
            m = self.ms[rd]
            local0 = self.local_start[rd]
            first = w * local0                      # the round's first stack slot = first projection of the scan
            mine_first = first + self.plan.rank * m
            if upload:
                for j in range(m):
                    d = fctx.dev_alloc(self.px * 4)
                    fctx.proj_h2d(self.h_raw.ptr + (local0 + j) * self.px * 4, d, self.det.n_row, self.det.n_col)
                    fctx.filter_to_stack(d, self.det, self.filter, self.d_stack, mine_first + j, self.layout)
                    fctx.dev_free(d)
            else:
                for done_m in range(0, m, 64):      # (a filter launch covers at most 256 projections)
                    n = min(64, m - done_m)
                    fctx.filter_to_stack_batch(self.d_raw + (local0 + done_m) * self.px * 4, self.px, n, self.det,
                                               self.filter, self.d_stack, mine_first + done_m, self.layout)
            filtered = torch.cuda.Event()
            filtered.record(self._filter_stream)
            with torch.cuda.stream(self._comm_stream):
                self._comm_stream.wait_event(filtered)
                out = self._torch_stack[first * slot_floats:(first + w * m) * slot_floats]
                mine = self._torch_stack[mine_first * slot_floats:(mine_first + m) * slot_floats]
                if ev_trace:
                    a = torch.cuda.Event(enable_timing=True)
                    a.record(self._comm_stream)
                self.dist.all_gather_into_tensor(out, mine)
                done = torch.cuda.Event(enable_timing=bool(ev_trace))
                done.record(self._comm_stream)
                if ev_trace:
                    comm_marks.append((a, done))
            # every round is backprojected as soon as it has arrived (backproject_stack cuts it into launches of at
            # most one batch)
            self._ext_stream.wait_event(done)
            self._backproject(first, w * m, download=upload and rd == self.rounds - 1)
        t_submitted = time.perf_counter()
        if ev_trace:
            t_ev1 = torch.cuda.Event(enable_timing=True)
            t_ev1.record(self._ext_stream)
            t_ev1.synchronize()
            busy = sum(a.elapsed_time(b) for a, b in comm_marks)
            ends = [t_ev0.elapsed_time(b) for _, b in comm_marks]
            if self.plan.rank == 0:
                print(f"[events] step {t_ev0.elapsed_time(t_ev1):.1f} ms; exchange busy {busy:.1f} ms over {len(comm_marks)} "
                      f"rounds of {[w * x for x in self.ms]}; round ends at {[round(e, 1) for e in ends[:4]]} ... {[round(e, 1) for e in ends[-3:]]}",
                      flush=True)
        if trace:
            t_end = time.perf_counter()
            print(f"[trace rank {self.plan.rank}] rounds submitted in {(t_submitted - t_begin) * 1e3:.1f} ms, "
                  f"last rounds + download {(t_end - t_submitted) * 1e3:.1f} ms, pool {ctx.stats()}", flush=True)

    def step_resident(self, timed: bool = False, overlap: bool = True):
        """raw projections already in HBM -> slab in HBM.  timed (sequential, for the stage breakdown):
        returns (filter, all-gather, backproject) milliseconds."""
        ctx = self.ctx
        if self.dist is not None and self.m and overlap and not timed:
            self._pipelined(upload=False)
            return None
        e0 = ctx.event() if timed else None
        ctx.volume_clear(self.d_vol, *self.slab_dims)
        if self.m:
            for rd in range(self.rounds):
                for done_m in range(0, self.ms[rd], 64):
                    n = min(64, self.ms[rd] - done_m)
                    local = self.local_start[rd] + done_m
                    ctx.filter_to_stack_batch(self.d_raw + local * self.px * 4, self.px, n, self.det, self.filter,
                                              self.d_stack, self.global_index(local), self.layout)
        elif self.my_count:
            ctx.filter_to_stack_batch(self.d_raw, self.px, self.my_count, self.det, self.filter, self.d_stack, self.lo,
                                      self.layout)
        e1 = ctx.event() if timed else None
        self._allgather()
        e2 = ctx.event() if timed else None
        self._backproject(0, self.n_proj)
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
            dropin.reconstruct(self.h_raw.ptr, self.n_proj, self.det, self.vol, self.h_slab.ptr, self.region,
                               roi=self.roi, device=self.device)
            return
        ctx = self.ctx
        if self.m:
            self._pipelined(upload=True)   # (ends with the overlapped backprojection + download of the last round)
            return
        else:
            ctx.volume_clear(self.d_vol, *self.slab_dims)
            for i in range(self.my_count):
                d = ctx.dev_alloc(self.px * 4)
                ctx.proj_h2d(self.h_raw.ptr + i * self.px * 4, d, self.det.n_row, self.det.n_col)
                ctx.filter_to_stack(d, self.det, self.filter, self.d_stack, self.lo + i, self.layout)
                ctx.dev_free(d)
            self._allgather()
            self._backproject(0, self.n_proj)
        ctx.vol_d2h(self.d_vol, self.h_slab.ptr, self.slab_dims[0] * self.slab_dims[1] * self.slab_dims[2])

    def slab(self) -> np.ndarray:
        return self.h_slab.array

    def close(self):
        self.ctx.sync()
        if self.fctx is not self.ctx:
            self.fctx.sync()
            self.fctx.close()
        self.ctx.filter_destroy(self.filter)
        self.ctx.volume_free(self.d_vol)
        if self.dist is None:
            self.ctx.stack_free(self.d_stack)
        if self.d_raw is not None:
            self.ctx.dev_free(self.d_raw)
        if self.h_raw is not None:
            self.h_raw.free()
        self.h_slab.free()
        self._torch_stack = None
