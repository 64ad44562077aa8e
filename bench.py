#!/usr/bin/env python
"""bench.py -- FDK hot path (weight -> ramp filter -> backprojection) on N B200s, one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config c1|c2|c3|c4|c5]

Default workload: BASELINE config 3 (1024^3 from 1440 x 2048^2), the configuration north_star quotes its target on;
it fits one GPU (24 GB raw + 24 GB filtered stack + 4 GB volume).

A "step" is one full reconstruction of the workload from its raw projections:
  value   device-resident: raw stack already in HBM; per step the fused weight+filter kernel fills the
          filtered stack and the batched backprojection kernel builds the volume (CUDA events on the
          library's compute stream).  GUPS = voxels x projections / seconds / 1e9.
  e2e     the same reconstruction through the reference-shaped per-projection loop of the C++ host layer
          (load -> weight -> filter -> backproject per projection, then copy_d2h): raw projections start
          in pinned HOST memory and the volume ends in pinned HOST memory, copies inside the timed region.
N > 1 (torchrun, one rank per GPU, process group over NCCL): the region is cut into N z-slabs; every rank uploads and
filters 1/N of the projections round by round and copies the detector-row band each peer needs straight into that
peer's stack over NVLink (paris_b200_group_*, csrc/group.cu); every rank backprojects all projections into its own
slab (no reduction) and downloads it into ONE shared host volume.  Strong scaling: the workload is the same for every N.

--impl reference times the reference's own OpenMP backend (oracle/_ref, built from /root/reference; the
plain-C port if that is absent) on the host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (detector n, projections, volume k, description)
    "c1": (256, 256, 128, "128^3 volume from 256 projections of 256^2 detector"),
    "c2": (1024, 720, 512, "512^3 volume from 720 projections of 1024^2 detector, full-scan FDK, float32"),
    "c3": (2048, 1440, 1024, "1024^3 volume from 1440 projections of 2048^2 detector"),
    # ROI configurations: the PARIS-derived natural volume, reconstructed inside a region of interest
    "c4": (2048, 2880, 1024, "ROI reconstruction (region_of_interest) 1024^3 sub-volume from 2880 projections of "
                             "2048^2 offset detector"),
    "c5": (2048, 2880, 2048, "2048^3 volume (32 GB float32) from 2880 projections of 2048^2, z-slabs across the GPUs"),
}


class _OracleApi:
    """The reference arm's geometry types and arithmetic: oracle/ only (that arm must not load the product)."""

    def __init__(self):
        import oracle
        self._impl = oracle.Reference() if oracle.have_ref() else oracle.Port()
        self.DetectorGeometry, self.VolumeGeometry, self.Roi = oracle.DetectorGeometry, oracle.VolumeGeometry, oracle.Roi
        self.calculate_volume_geometry, self.apply_roi = self._impl.calculate_volume_geometry, self._impl.apply_roi


def geometry(cfg: str, capi=None):
    """(detector, FULL volume geometry, projections, roi or None, region (x, y, z)).  `capi`: the module or object
    that provides the geometry structs and calculate_volume_geometry / apply_roi (default: the product's C ABI)."""
    if capi is None:
        from paris_b200 import capi
    n, n_proj, k, _ = CONFIGS[cfg]
    l_px = 0.2 * 1024 / n  # 204.8 mm detector
    delta_s = 100.0 if cfg == "c4" else 0.0   # offset detector (pixels)
    det = capi.DetectorGeometry(n, n, l_px, l_px, delta_s, 0.0, 500.0, 500.0, 360.0 / n_proj)
    nat = capi.calculate_volume_geometry(det)
    if cfg in ("c4", "c5"):
        # a centred k^3 box of the natural volume; apply_roi: dim = x2 - x1, + 1 iff x1 == 0 (src/geometry.cpp:86-130)
        def span(full):
            lo = (full - k) // 2
            return lo, (lo + k - 1) if lo == 0 else (lo + k)
        (x1, x2), (y1, y2), (z1, z2) = span(nat.dim_x), span(nat.dim_y), span(nat.dim_z)
        roi = capi.Roi(x1, x2, y1, y2, z1, z2)
        reg = capi.apply_roi(nat, roi)
        assert (reg.dim_x, reg.dim_y, reg.dim_z) == (k, k, k), (reg.dim_x, reg.dim_y, reg.dim_z)
        return det, nat, n_proj, roi, (k, k, k)
    # "K^3 from a (2K)^2 detector": the natural full-FOV volume sampled with K^3 larger voxels
    f32 = np.float32
    vol = capi.VolumeGeometry(k, k, k, f32(nat.l_vx_x * nat.dim_x / k), f32(nat.l_vx_y * nat.dim_y / k),
                              f32(nat.l_vx_z * nat.dim_z / k))
    return det, vol, n_proj, None, (k, k, k)


def ellipsoids(det, phantom=None):
    if phantom is None:
        from paris_b200 import phantom
    r = 0.9 * phantom.fov_radius(det.n_row, det.l_px_row, det.delta_s, det.d_so, det.d_od)
    return phantom.scaled_ellipsoids(phantom.SHEPP_LOGAN_3D, r)


class ClockSampler:
    """SM clock / throttle reasons DURING the timed region (B200_PROFILING.md recipe), read through NVML
    from a background thread.  (An `nvidia-smi -lms 100` child process was measured to stall this process's
    CUDA calls for hundreds of milliseconds per step; three light NVML queries do not.)"""

    # NB: NVML queries contend with CUDA submissions inside the driver.  Sampling every 100 ms is harmless for
    # the device-resident steps (a few dozen launches per step) but was measured to stretch the end-to-end
    # steps (thousands of API calls each) from 139 ms to as much as 1 s -- so that phase is sampled at 500 ms.
    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, device: int, interval_ms: int = 100):
        self.device, self.interval = device, interval_ms / 1e3
        self.samples, self.power, self.reasons = [], [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None

    def _sample(self):
        import pynvml
        h = self._handle
        self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
        self.power.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1e3)
        mask = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        for name, bit in self.REASONS.items():
            if mask & bit:
                self.reasons.add(name)

    def _run(self):
        try:
            while not self._stop.is_set():
                self._sample()
                self._stop.wait(self.interval)
        except Exception as e:  # report that instead of clocks
            self.error = repr(e)

    def start(self):
        """NVML is initialised HERE, synchronously (importing and initialising it takes longer than a short timed
        region); the thread only samples."""
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            index = self.device
            if visible:
                try:
                    index = int(visible.split(",")[self.device])
                except (ValueError, IndexError):
                    pass
            self._handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._handle, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # no NVML: report that instead of clocks
            self.error = repr(e)
            return
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        if self._thread is None:
            return out
        self._stop.set()
        self._thread.join(timeout=5)
        if self.samples:
            # "under load": the upper half of the samples (the sampler also sees the gaps between steps)
            busy = sorted(self.samples)[len(self.samples) // 2:]
            out.update(sm_mhz=float(np.median(busy)), sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons),
                       samples=len(self.samples), power_w_max=float(max(self.power)))
        if hasattr(self, "error"):
            out["error"] = self.error
        return out


def ncu_record(kernel: str):
    """The newest committed ncu summary of `kernel` (profiles/*_ncu_summary.json: one --set full capture each),
    or None.  A record may carry "capture": {"what": text, "algorithmic_bytes": B} describing the launch that was
    captured, so that DRAM traffic is compared with the algorithmic bytes of THAT launch."""
    import glob
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_ncu_summary.json")), reverse=True):
        try:
            rec = json.load(open(path)).get(kernel)
        except (OSError, ValueError):
            continue
        if rec:
            rec = dict(rec)
            rec["_file"] = os.path.basename(path)
            return rec
    return None


def ncu_traffic(rec):
    """dram__bytes_read.sum + dram__bytes_write.sum of the captured launch, in bytes."""
    if not rec:
        return None
    total = 0.0
    for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        unit = rec.get(key + "__unit", "byte").lower()
        scale = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1.0)
        total += float(rec.get(key, 0.0)) * scale
    return total


def traffic_note(rec):
    if not rec:
        return "no ncu capture committed"
    cap = rec.get("capture") or {}
    note = f"DRAM bytes of the launch captured in profiles/{rec['_file']}"
    if cap.get("what"):
        note += f" ({cap['what']})"
    if cap.get("algorithmic_bytes"):
        note += f"; algorithmic HBM bytes of that launch = {float(cap['algorithmic_bytes']):.3e}"
    return note


def gather_peak():
    """The measured ceiling of the backprojection's shared-memory gather (scripts/microbench/gather_peak.cu, committed
    output profiles/r2_gather_peak.json): GUPS of the kernel's own LDS.32 pattern with nothing else in the way."""
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "r2_gather_peak.json")))
        return {"gups": rec["gather_gups"], "source": "profiles/r2_gather_peak.json"}
    except (OSError, ValueError, KeyError):
        return None


def measured_peaks() -> dict:
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        return {}


# ------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's OpenMP backend (or the plain-C port) on a bounded sample of the workload
# ------------------------------------------------------------------------------------------------------------------

def host_threads() -> int:
    """Host cores this process may use (its affinity mask), NOT what OMP_NUM_THREADS says: torchrun exports
    OMP_NUM_THREADS=1 to every rank, which made the round-1 reference arm single-threaded at N > 1."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_reconstruct_sample(cfg: str, budget_s: float, fetch=None):
    """Times weight+filter+backproject of the reference on `p` projections of the workload (full region), with
    every host core.  Imports oracle/ only.  Returns (gups_total, gups_backprojection, cores, kind, description, seconds)."""
    import oracle
    import oracle.phantom as ophantom
    api = _OracleApi()
    det, vol, n_proj, roi, region = geometry(cfg, api)
    odet, ovol, oroi = det, vol, roi
    if oracle.have_ref():
        impl, kind = oracle.Reference(), "reference"
    else:
        impl, kind = oracle.Port(), "port"
    impl.set_num_threads(host_threads())
    cores = impl.num_threads()
    shape = (region[2], region[1], region[0])
    voxels = region[0] * region[1] * region[2]

    def raw(count, stride):
        if fetch is not None:
            return fetch(count, stride)
        ang = np.float32(det.delta_phi) * (np.arange(count, dtype=np.float32) * np.float32(stride))
        return ophantom.project(ellipsoids(det, ophantom), det.n_row, det.n_col, det.l_px_row, det.l_px_col, det.delta_s,
                                det.delta_t, det.d_so, det.d_od, ang.astype(np.float64))

    # calibrate on 2 projections (the first builds FFT plans / statics and is excluded by the callee)
    stride = max(1, n_proj // 16)
    _, t = impl.reconstruct(raw(2, stride), shape, odet, ovol, roi=oroi, idx_stride=stride)
    per_proj = max(sum(t), 1e-4)
    count = int(max(2, min(64, budget_s / per_proj))) + 1
    stride = max(1, n_proj // count)
    t0 = time.perf_counter()
    _, t = impl.reconstruct(raw(count, stride), shape, odet, ovol, roi=oroi, idx_stride=stride)
    wall = time.perf_counter() - t0
    timed = count - 1
    gups_total = voxels * timed / sum(t) / 1e9
    gups_bp = voxels * timed / t[2] / 1e9
    desc = (f"{timed} of {n_proj} projections (every {stride}th) of {CONFIGS[cfg][3]}, full region, "
            f"weight {t[0]:.2f}s + filter {t[1]:.2f}s + backproject {t[2]:.2f}s")
    return gups_total, gups_bp, cores, kind, desc, wall


def run_reference(args, rank: int):
    """The reference's own CPU implementation (oracle/_ref = /root/reference compiled unmodified; the plain-C port
    where that is absent) on every host core.  ONE bounded sample (a 2-projection calibration run before it is the
    warm-up: FFT plans, function-local statics): per-update cost does not depend on the projection, so repeating
    the sample --steps times would only burn lease time.  Loads nothing of the product."""
    if rank != 0:
        return
    assert "paris_b200" not in sys.modules, "the reference arm must not load the product"
    api = _OracleApi()
    det, vol, n_proj, _roi, region = geometry(args.config, api)
    g_total, g_bp, cores, kind, desc, wall = cpu_reconstruct_sample(args.config, budget_s=args.cpu_budget)
    assert "paris_b200" not in sys.modules
    value = float(g_total)
    updates = region[0] * region[1] * region[2] * n_proj
    line = {
        "impl": "reference", "metric": "fdk_reconstruction_gups", "value": value, "unit": "GUPS",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": updates / value / 1e6, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": CONFIGS[args.config][3],
                   "note": "one bounded sample (after a 2-projection warm-up run) stands for every step; ms_per_step "
                           "extrapolated from it"},
        "cpu_baseline": {"value": value, "unit": "GUPS", "cores": cores, "kind": kind, "sample": desc,
                         "backprojection_only_gups": g_bp, "sample_wall_s": wall},
        "e2e": {"value": value, "unit": "GUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------------------------

class SharedHostVolume:
    """The host volume every member downloads its slabs into (the sink's job, /root/reference/src/sink.cpp:76-81: the
    volume is assembled by writing every slab at its offset).  One POSIX shared-memory segment, page-locked by every
    process; falls back to a pinned buffer of the member's own where that is not possible (then `kind` says so)."""

    def __init__(self, capi, dist, rank, world, voxels, member, shared, region_x):
        self.capi, self.kind, self.shm, self.registered = capi, "per-rank pinned buffers", None, False
        self.ptr = None
        ok = True
        if shared:   # (volumes beyond 8 GB: every member keeps its slabs in a buffer of its own)
            from multiprocessing import shared_memory
            name = f"paris_b200_bench_{os.environ.get('MASTER_PORT', '0')}_{os.getppid()}"
            try:
                if rank == 0:
                    try:
                        shared_memory.SharedMemory(name=name).unlink()
                    except FileNotFoundError:
                        pass
                    self.shm = shared_memory.SharedMemory(name=name, create=True, size=voxels * 4)
            except OSError:
                ok = False
            dist.barrier()
            try:
                if rank != 0 and ok:
                    self.shm = shared_memory.SharedMemory(name=name)
                    try:   # (rank 0 owns the segment: the other ranks' resource trackers must not unlink it again)
                        from multiprocessing import resource_tracker
                        resource_tracker.unregister(self.shm._name, "shared_memory")
                    except Exception:
                        pass
                if self.shm is not None:
                    self.array = np.ndarray((voxels,), np.float32, buffer=self.shm.buf)
                    capi.check(capi.lib().paris_b200_host_register(self.array.ctypes.data, voxels * 4))
                    self.registered = True
            except (OSError, capi.Error):
                ok = False
            import torch
            flag = torch.tensor([1.0 if (ok and self.registered) else 0.0], device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if flag.item() > 0:
                self.kind = "one shared page-locked volume (POSIX shm), slabs written at their offsets"
                self.ptr = self.array.ctypes.data + member.host_offset_bytes(region_x)
                member.group.set_host_row(region_x)    # the member's box sits inside the region-wide volume
            else:
                self.kind = "per-rank pinned buffers (the shared segment could not be page-locked by every rank)"
        elif world > 1:
            self.kind = "per-rank pinned buffers (volume larger than 8 GB)"
        if self.ptr is None:
            self.ptr = member.alloc_host_slabs().ptr

    def close(self, rank):
        if self.registered:
            self.capi.lib().paris_b200_host_unregister(self.array.ctypes.data)
        self.array = None
        if self.shm is not None:
            self.shm.close()
            if rank == 0:
                self.shm.unlink()


def auto_x_parts(world: int, dims, spm: int, full_region: bool) -> int:
    """Parts along x of the boxes the GPUs own (see the comment where run_b200 calls it)."""
    x_parts = 1
    if world >= 4 and world % 2 == 0 and spm == 1 and full_region and dims[0] // (world // 2) >= 64:
        x_parts = world // 2
    while x_parts < world and world % (2 * x_parts) == 0 and dims[2] * x_parts // world < 128 * spm:
        x_parts *= 2
    return x_parts


def run_b200(args, rank: int, world: int, local_rank: int):
    from paris_b200 import capi, dropin, phantom
    from paris_b200.multi import GroupMember
    from paris_b200.pipeline import angle_sin_cos

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(local_rank)
        # one rank per GPU over NCCL: handles, barriers and the max-over-ranks of the timings travel through it; the
        # filtered projections themselves go through peer memory (csrc/group.cu)
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist = dist_mod

    det, vol, n_proj, roi, dims = geometry(args.config)
    n = det.n_row
    px = n * n
    voxels = dims[0] * dims[1] * dims[2]
    updates = voxels * n_proj
    spm = args.slabs_per_gpu if args.slabs_per_gpu else (2 if args.config == "c5" else 1)

    # Which box of the region a GPU owns.  Plain z-slabs are not equally expensive: the outermost 128 slices cost 11 %
    # more than inner ones (profiles/r2_slab_times_c3.json: their near-source voxels project beyond the detector's
    # top and bottom rows, so their tiles run the boundary path), and the step is as long as its slowest slab.  From
    # four GPUs on the region is therefore cut into TWO mirror z-runs, each of which holds one outermost block, times
    # world/2 parts along x (measured at N=8: config 3 174.9 instead of 181.8 ms, config 2 unchanged; at N=4 the
    # four boxes are mirror images of each other).  It also keeps z-runs of >= 128 slices and with them the 8x8x128
    # tiles on volumes with few slices per GPU (config 2 at N=8: 64).
    x_parts = args.x_parts if args.x_parts else auto_x_parts(world, dims, spm, roi is None)
    shared_volume = world > 1 and voxels * 4 <= (8 << 30)
    u16 = args.samples == "u16"
    member = GroupMember(local_rank, rank, world, det, vol, n_proj, roi=roi, slabs_per_member=spm, x_parts=x_parts,
                         whole_projections=bool(args.whole_projections),
                         sample_type=capi.SAMPLES_U16 if u16 else capi.SAMPLES_F32,
                         exchange=capi.EXCHANGE_KERNEL if args.exchange == "kernel" else capi.EXCHANGE_COPY_ENGINE)
    if dist is not None:
        handles = [None] * world
        dist.all_gather_object(handles, member.export())
        member.connect(handles)
    ctx, info = member.ctx, member.info
    # --samples u16: the phantom's line integrals as 16-bit detector counts (the longest chord stays below 65536)
    member.generate_inputs(ellipsoids(det), counts_scale=phantom.counts_scale(ellipsoids(det)) if u16 else None)
    sc = np.array([angle_sin_cos(i, det) for i in range(n_proj)], dtype=np.float32)
    my_slices = info.z_count
    slab_dims = (info.x_count, dims[1], my_slices)
    box_roi = member.box_roi(roi)          # (the ROI that shifts voxel indices to this member's columns)
    # stage kernels on their own (roofline legs): this member's share through the filter, all projections through the
    # backprojection into this member's slices, on ONE stream with events in between
    filt = ctx.filter_create(capi.filter_size(n), float(det.l_px_row))
    stage_vol = info.d_first_slab if info.slabs == 1 else ctx.volume_alloc(*slab_dims)

    def staged_step():
        e0 = ctx.event()
        local = 0
        for first, count in member.runs:
            for done in range(0, count, 64):
                c = min(64, count - done)
                (ctx.filter_to_stack_batch_u16 if u16 else ctx.filter_to_stack_batch)(
                    member.d_raw + (local + done) * member.proj_bytes, px, c, det, filt, info.d_stack, first + done, info.layout)
            local += count
        e1 = ctx.event()
        ctx.volume_clear(stage_vol, *slab_dims)
        ctx.backproject_stack(info.d_stack, 0, n_proj, sc[:, 0], sc[:, 1], stage_vol, slab_dims, info.z_first, det, vol,
                              roi=box_roi, layout=info.layout)
        e2 = ctx.event()
        tf = ctx.elapsed_ms(e0, e1, destroy=False)
        tb = ctx.elapsed_ms(e1, e2, destroy=False)
        for e in (e0, e1, e2):
            capi.check(capi.lib().paris_b200_event_destroy(e))
        return tf, tb

    sampler = ClockSampler(local_rank, args.clock_interval_ms)
    barrier = (lambda: dist.barrier()) if dist is not None else (lambda: None)

    # ---- device-resident steps -------------------------------------------------------------------------------
    t_filter = t_bp = 0.0
    if world == 1:
        # one GPU: filter, then backprojection, back to back on one stream -- the timed steps themselves give the
        # per-kernel durations of the roofline objects
        for _ in range(args.warmup):
            staged_step()
        ctx.sync()
        sampler.start()
        launches0 = member.launch_count()
        t0 = time.perf_counter()
        e_start = ctx.event()
        for _ in range(args.steps):
            tf, tb = staged_step()
            t_filter += tf
            t_bp += tb
        e_stop = ctx.event()
        ms_total = ctx.elapsed_ms(e_start, e_stop)
        ctx.sync()
    else:
        # N GPUs: the pipelined group step (upload-free: raw projections resident) -- filter, exchange over peer
        # memory and backprojection overlap; timed with events on the backprojection stream, max over ranks
        for _ in range(args.warmup):
            member.step_resident()
        barrier()
        sampler.start()
        launches0 = member.launch_count()
        t0 = time.perf_counter()
        e_start = ctx.event()
        for _ in range(args.steps):
            member.step_resident()
        e_stop = ctx.event()
        ms_total = ctx.elapsed_ms(e_start, e_stop)
        barrier()
    wall_resident = time.perf_counter() - t0
    launches = member.launch_count() - launches0
    clocks = sampler.stop()
    kernel_info = ctx.bp_kernel_info()   # (the instantiation the resident steps ran, before the e2e path's chunked tail)

    parity = None
    if world > 1:
        # the volumes, not just the speed: two 4-slice bands of this member's slab recomputed from its own stack by the
        # exact kernel (the reference's arithmetic operation for operation, tests/test_gpu_parity.py)
        c = float(n_proj) / (8.0 * np.pi) * (phantom.counts_scale(ellipsoids(det)) if u16 else 1.0)
        worst = [0.0, 0.0]
        first_dz = member.plan.slab_dz if info.slabs > 1 else my_slices   # (bands of the member's FIRST slab)
        if True:
            for z in sorted({0, max(0, first_dz // 2 - 2), max(0, first_dz - 4)}):
                dz = min(4, first_dz - z)
                got = member.device_slab(z, dz)
                ctx.set_option("bp_kernel", 1)
                v = ctx.volume_alloc(info.x_count, dims[1], dz)
                ctx.backproject_stack(info.d_stack, 0, n_proj, sc[:, 0], sc[:, 1], v, (info.x_count, dims[1], dz), info.z_first + z,
                                      det, vol, roi=box_roi, layout=info.layout)
                want = np.empty_like(got)
                ctx.vol_d2h(v, want, want.size)
                ctx.volume_free(v)
                ctx.set_option("bp_kernel", 0)
                d = got.astype(np.float64) - want
                worst[0] = max(worst[0], float(np.abs(d).max() / c))
                worst[1] = max(worst[1], float(np.sqrt(np.mean(d * d)) / c))
                if not np.isfinite(got).all():
                    worst = [float("inf"), float("inf")]
        parity = worst
        # stage breakdown (diagnostic, outside the timed region): the same kernels without overlap
        for _ in range(min(2, args.steps)):
            tf, tb = staged_step()
            t_filter += tf * args.steps / min(2, args.steps)
            t_bp += tb * args.steps / min(2, args.steps)
        ctx.sync()
        barrier()

    # ---- end-to-end steps: pinned host projections -> the host volume --------------------------------------------------
    host_vol = SharedHostVolume(capi, dist, rank, world, voxels, member, shared_volume, dims[0]) if world > 1 else None
    if world == 1:
        member.alloc_host_slabs()

    def e2e_step():
        if world == 1 and u16:
            member.step_e2e()        # (the drop-in loop takes the reference's float projections; counts go through the group)
        elif world == 1:
            # the reference-shaped per-projection loop in C++ (paris_b200/cpp/pipeline.cpp: reconstruct_task)
            dropin.reconstruct(member.h_raw.ptr, n_proj, det, vol, member.h_slabs.ptr, dims, roi=roi, device=local_rank)
        else:
            member.step_e2e(host_vol.ptr)

    for _ in range(max(1, args.warmup)):
        e2e_step()
    barrier()
    sampler_e2e = ClockSampler(local_rank, 500)
    sampler_e2e.start()
    e2e_ms = []
    for _ in range(args.steps):
        barrier()
        t1 = time.perf_counter()
        e2e_step()
        barrier()    # (the volume is complete when every member's slabs are in it)
        e2e_ms.append((time.perf_counter() - t1) * 1e3)
    clocks_e2e = sampler_e2e.stop()
    clocks["reasons"] = sorted(set(clocks["reasons"]) | set(clocks_e2e["reasons"]))
    clocks["e2e_phase"] = {k: clocks_e2e.get(k) for k in ("sm_mhz", "samples", "power_w_max")}
    pushed = member.group.info().bytes_pushed
    memops = member.group.info().memops

    ms_step = ms_total / args.steps
    e2e_step_ms = float(np.mean(e2e_ms))
    if dist is not None:
        import torch
        t = torch.tensor([ms_step, e2e_step_ms, t_filter, t_bp, parity[0], parity[1]], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step, e2e_step_ms, t_filter, t_bp, parity[0], parity[1] = [float(x) for x in t.tolist()]

    if rank == 0:
        peaks = measured_peaks()
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured" if "hbm_gbs" in peaks else "fallback"
        sm_mhz = clocks.get("sm_mhz") or float(peaks.get("sm_max_mhz", 1965.0))
        # backprojection: 4 point-fetched float samples per voxel update from shared memory
        bp_s = t_bp / args.steps / 1e3
        filt_s = t_filter / args.steps / 1e3
        my_updates = my_slices * info.x_count * dims[1] * n_proj
        bp_gbs = 16.0 * my_updates / bp_s / 1e9
        smem_peak = 148 * 128 * sm_mhz * 1e6 / 1e9
        filt_gbs = (4.0 + member.sample_bytes) * px * member.my_count / filt_s / 1e9
        ncu_bp, ncu_filt = ncu_record("backprojection"), ncu_record("fused")
        gp = gather_peak()
        bands = [(member.plan.band_lo[k], member.plan.band_hi[k]) for k in range(world)]
        line = {
            "metric": "fdk_reconstruction_gups", "value": updates / (ms_step / 1e3) / 1e9, "unit": "GUPS",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": CONFIGS[args.config][3],
                       "parallelism": (f"z-slab x{world}" if x_parts == 1 else f"{world // x_parts} z-runs x {x_parts} x-parts")
                                      + (f", {spm} slabs per GPU" if spm > 1 else ""),
                       "l2": "inputs larger than L2 (raw + filtered stack = 48 GB at config 3)",
                       "bp_batch": 256,
                       "stack_layout": "transposed, v fastest, " + ("parity-split" if info.layout else "plain"),
                       "exchange": ("none" if world == 1 else
                                    ("whole projections" if args.whole_projections else "detector-row bands") + " into peer memory, "
                                    + ("copy kernel" if args.exchange == "kernel" else "copy engines")
                                    + (", stream memory operations" if memops else ", one-word kernels") + " as arrival flags")},
            "backprojection_gups": my_updates * world / bp_s / 1e9,
            "stage_ms": {"filter": filt_s * 1e3, "backproject": bp_s * 1e3,
                         "note": "the two kernels on one stream without overlap (at N>1 measured next to the timed region, "
                                 "whose step overlaps upload, filter, exchange and backprojection)"},
            "roofline": {"kernel": kernel_info["last"], "bound": "smem", "achieved": bp_gbs, "peak": smem_peak,
                         "unit": "GB/s", "frac": bp_gbs / smem_peak, "traffic": ncu_traffic(ncu_bp),
                         "traffic_note": traffic_note(ncu_bp),
                         "launches_by_kernel": {"tma": kernel_info["tma_launches"], "exact": kernel_info["exact_launches"]},
                         "measured_gather_peak": gp,
                         "frac_of_measured_gather_peak": (my_updates / bp_s / 1e9 / gp["gups"]["dv_c2c3_distribution"]
                                                          if gp and args.config in ("c1", "c2", "c3") else None),
                         "pipe_busy_ncu_pct": (ncu_bp or {}).get(
                             "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed"),
                         "issue_active_ncu_pct": (ncu_bp or {}).get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                         "pipe_busy_note": "ncu: share of cycles the L1/shared-memory data pipe is busy for this kernel "
                                           "(algorithmic bytes + bank conflicts + table broadcasts), profiles/",
                         "note": "16 B of shared-memory sample fetches per voxel update; peak = 148 SMs x 128 B/clk x "
                                 f"{sm_mhz:.0f} MHz (SM clock sampled during the run); not an HBM- or tensor-bound kernel.  "
                                 "measured_gather_peak: the same LDS.32 pattern alone (scripts/microbench/gather_peak.cu): "
                                 "lanes along z advance 0.87..1.17 words per lane, every load whose 32 rows span more "
                                 "than 32 banks costs two wavefronts"},
            "roofline_filter": {"kernel": "filter_kernel", "bound": "hbm", "achieved": filt_gbs, "peak": hbm_peak,
                                "unit": "GB/s", "frac": filt_gbs / hbm_peak, "traffic": ncu_traffic(ncu_filt),
                                "traffic_note": traffic_note(ncu_filt),
                                "pipe_busy_ncu_pct": (ncu_filt or {}).get(
                                    "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed"),
                                "note": f"8 B per detector pixel; peak {peak_src} (MEASURED_PEAKS.json hbm_gbs); ~110 flop and "
                                        "~30 shared-memory accesses per pixel keep it on the FP32/shared-memory side of the ridge"},
            "e2e": {"value": updates / (e2e_step_ms / 1e3) / 1e9, "unit": "GUPS", "seconds": e2e_step_ms / 1e3,
                    "h2d_bytes_per_step": member.proj_bytes * n_proj, "d2h_bytes_per_step": 4 * voxels,
                    "samples": "u16 detector counts, widened by the filter kernel" if u16 else "f32",
                    "ms_steps": [round(x, 2) for x in e2e_ms],
                    "path": ("paris_b200_dropin_reconstruct: per-projection load/weight/filter/backproject + copy_d2h"
                             if world == 1 and not u16 else
                             "paris_b200_group_reconstruct per rank: upload + filter of 1/N of the projections round by round, "
                             "exchange over peer memory, backprojection of all projections into the rank's slabs, download"),
                    "host_volume": ("one pinned buffer" if world == 1 else host_vol.kind)},
            "gpu_launches": int(launches), "clocks": clocks,
            "wall_s_resident": wall_resident,
        }
        if world > 1:
            group_steps = args.warmup + args.steps + max(1, args.warmup) + args.steps
            line["exchange"] = {"bytes_pushed_per_step_rank0": int(pushed // group_steps),
                                "all_gather_bytes_per_step_rank0": int((world - 1) * member.my_count * px * 4),
                                "band_rows_per_rank": [hi - lo for lo, hi in bands], "detector_rows": int(det.n_col)}
            line["parity_check"] = {"max": parity[0], "rmse": parity[1], "ok": bool(parity[0] <= 1e-4 and parity[1] <= 1e-5),
                                    "what": "every rank: three 4-slice bands of its (first) slab (first, middle, last slices) after "
                                            "the timed steps against the exact kernel on the rank's own gathered stack; "
                                            "max over ranks, in units of the phantom contrast"}
        if world == 1 and not args.no_cpu_baseline:
            try:
                g_total, g_bp, cores, kind, desc, _ = cpu_reconstruct_sample(args.config, budget_s=args.cpu_budget,
                                                                           fetch=member.host_sample)
                line["cpu_baseline"] = {"value": g_total, "unit": "GUPS", "cores": cores, "kind": kind, "sample": desc,
                                        "backprojection_only_gups": g_bp}
            except Exception as e:  # the CPU checker must never take the GPU number down with it
                line["cpu_baseline"] = {"value": None, "unit": "GUPS", "cores": None, "kind": "unavailable",
                                        "sample": repr(e)}
        print(json.dumps(line), flush=True)

    ctx.filter_destroy(filt)
    if info.slabs != 1:
        ctx.volume_free(stage_vol)
    if dist is not None:
        dist.barrier()        # nobody tears its stack down while a peer may still push into it
    if host_vol is not None:
        host_vol.close(rank)
    member.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS))
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU work per reference sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--clock-interval-ms", type=int, default=100, help="NVML sampling period")
    ap.add_argument("--slabs-per-gpu", type=int, default=0, help="z-slabs every GPU streams (0: 2 for config 5, else 1)")
    ap.add_argument("--x-parts", type=int, default=0, help="parts along x (0: automatic -- from 4 GPUs on two mirror z-runs x N/2 x-parts for full regions)")
    ap.add_argument("--exchange", default="copy-engine", choices=["copy-engine", "kernel"])
    ap.add_argument("--samples", default="f32", choices=["f32", "u16"],
                    help="raw projections as floats (what the reference's reader hands on) or as 16-bit detector counts")
    ap.add_argument("--whole-projections", action="store_true", help="exchange every detector row (an all-gather)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            sys.exit("bench.py --gpus N with N > 1 must be launched with torchrun (one rank per GPU)")
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
