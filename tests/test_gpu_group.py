"""The multi-GPU step behind the C ABI (paris_b200_group_*, csrc/group.cu).

On ONE GPU (what the driver's test box has) the whole machinery still runs: a group of one member, a member streaming
several slabs through two buffers, and groups of two and three members that all live on device 0 of this process --
they exchange their detector-row bands through each other's device pointers and wait on each other's arrival flags
exactly as members on different GPUs do (begin() only enqueues, so one host thread begins every member and then ends
them).  Every variant must reproduce the one-piece reconstruction BIT FOR BIT: rounds keep the projection order, tiles
are anchored globally, and a member's stack is NaN outside what it filtered itself and the band it received, so a band
chosen too small cannot go unnoticed.  With two or more GPUs the same is run with one process per GPU over CUDA IPC."""
import multiprocessing as mp
import os

import numpy as np
import pytest

from paris_b200 import capi, phantom
from paris_b200.multi import GroupMember
from paris_b200.pipeline import angle_sin_cos

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


def _case(k=64, extra_z=3, n=128, n_proj=48):
    det = capi.DetectorGeometry(n, n, 0.4, 0.4, 0.0, 0.0, 500.0, 500.0, 360.0 / n_proj)
    nat = capi.calculate_volume_geometry(det)
    f32 = np.float32
    vol = capi.VolumeGeometry(k, k, k + extra_z, f32(nat.l_vx_x * nat.dim_x / k), f32(nat.l_vx_y * nat.dim_y / k),
                              f32(nat.l_vx_z * nat.dim_z / (k + extra_z)))
    return det, vol, n_proj


def _ellipsoids(det):
    return phantom.scaled_ellipsoids(phantom.SHEPP_LOGAN_3D,
                                     0.9 * phantom.fov_radius(det.n_row, det.l_px_row, det.delta_s, det.d_so, det.d_od))


def _one_piece(ctx, det, vol, n_proj, roi=None, region=None):
    """the reference result of this file: filter_to_stack_batch + backproject_stack on one context"""
    region = region or (vol.dim_x, vol.dim_y, vol.dim_z)
    px = det.n_row * det.n_col
    raw = ctx.dev_alloc(n_proj * px * 4)
    ctx.phantom_project(_ellipsoids(det), det, 0, n_proj, raw)
    stack = ctx.stack_alloc(det.n_row, det.n_col, n_proj)
    filt = ctx.filter_create(capi.filter_size(det.n_row), float(det.l_px_row))
    layout = capi.choose_stack_layout(det, vol)
    for first in range(0, n_proj, 64):
        cnt = min(64, n_proj - first)
        ctx.filter_to_stack_batch(raw + first * px * 4, px, cnt, det, filt, stack, first, layout)
    sc = np.array([angle_sin_cos(i, det) for i in range(n_proj)], dtype=np.float32)
    ctx.set_option("bp_kernel", 2)
    v = ctx.volume_alloc(*region)
    ctx.backproject_stack(stack, 0, n_proj, sc[:, 0], sc[:, 1], v, region, 0, det, vol, roi=roi, layout=layout)
    out = np.empty((region[2], region[1], region[0]), np.float32)
    ctx.vol_d2h(v, out, out.size)
    ctx.set_option("bp_kernel", 0)
    ctx.volume_free(v)
    ctx.filter_destroy(filt)
    ctx.stack_free(stack)
    ctx.dev_free(raw)
    return out


def _poison(member):
    """NaN in every slot of the member's stack: whatever it does not filter itself or receive stays NaN"""
    info = member.info
    slot_bytes, _ = capi.stack_slot_bytes(member.det.n_row, member.det.n_col)
    nan = np.full(slot_bytes // 4, np.nan, np.float32)
    for i in range(member.n_proj):
        member.ctx.proj_h2d(nan, info.d_stack + i * slot_bytes, slot_bytes // 4, 1)
    member.ctx.sync()


def _run_in_process(world, det, vol, n_proj, roi=None, e2e=True, counts_scale=None, **options):
    """`world` members on device 0 of this process; returns the assembled region"""
    if options.get("x_parts", 1) > 1 or e2e:
        # every member downloads its box straight into ONE region-wide host volume
        region_x = capi.apply_roi(vol, roi).dim_x if roi is not None else vol.dim_x
        options = dict(options, host_row_floats=region_x)
    members = [GroupMember(0, r, world, det, vol, n_proj, roi=roi, **options) for r in range(world)]
    handles = [m.export() for m in members]
    for m in members:
        m.connect(handles)
        m.generate_inputs(_ellipsoids(det), counts_scale=counts_scale)
        _poison(m)
    info0 = members[0].info
    region = capi.PinnedArray((info0.region_z, info0.region_y, info0.region_x))
    region.array[...] = np.nan
    for step in range(2):                       # a second step exercises the write-after-read guards
        if e2e:
            for m in members:
                m.begin_e2e(region.ptr + m.host_offset_bytes(info0.region_x))
        else:
            for m in members:
                m.begin_resident()
        for m in members:
            m.end()
    if not e2e:
        for m in members:
            assert m.info.slabs == 1
            region.array[m.info.z_first:m.info.z_first + m.info.z_count, :,
                         m.info.x_first:m.info.x_first + m.info.x_count] = m.device_slab(0, m.info.z_count)
    out = region.array.copy()
    stats = [(m.group.info().bytes_pushed, m.group.info().memops, m.info.band_lo, m.info.band_hi) for m in members]
    for m in members:
        m.close()
    region.free()
    return out, stats


def test_group_of_one_equals_the_one_piece_path(ctx):
    det, vol, n_proj = _case()
    want = _one_piece(ctx, det, vol, n_proj)
    got, _ = _run_in_process(1, det, vol, n_proj, first_round=8, max_round=16)
    assert np.array_equal(got, want)
    got, _ = _run_in_process(1, det, vol, n_proj, e2e=False, first_round=8, max_round=16)
    assert np.array_equal(got, want)


def test_one_member_streams_several_slabs_through_two_buffers(ctx):
    det, vol, n_proj = _case()
    want = _one_piece(ctx, det, vol, n_proj)
    got, _ = _run_in_process(1, det, vol, n_proj, slabs_per_member=5, stream_slabs=True, first_round=16, max_round=16)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("exchange", [capi.EXCHANGE_COPY_ENGINE, capi.EXCHANGE_KERNEL])
@pytest.mark.parametrize("world,spm", [(2, 1), (3, 2)])
def test_members_sharing_one_gpu_exchange_bands_and_equal_one_piece(ctx, world, spm, exchange):
    det, vol, n_proj = _case()
    want = _one_piece(ctx, det, vol, n_proj)
    got, stats = _run_in_process(world, det, vol, n_proj, slabs_per_member=spm, stream_slabs=spm > 1, first_round=8,
                                 max_round=16, exchange=exchange)
    assert np.isfinite(got).all(), "a voxel read a detector row its member never received"
    assert np.array_equal(got, want)
    slot_bytes, _ = capi.stack_slot_bytes(det.n_row, det.n_col)
    for pushed, memops, lo, hi in stats:
        # two steps; per step a member pushes its share (n_proj / world slots) to world - 1 peers, bands only
        assert 0 < pushed < 2 * (world - 1) * (n_proj / world + 1) * slot_bytes
        print(f"pushed {pushed / 1e6:.1f} MB, band [{lo}, {hi}), memops {memops}")


def test_roi_region_offset_detector_natural_volume(ctx):
    """plain stack layout, ROI with odd offsets (STRADDLE tiles), detector shifted: three members on one GPU"""
    n, n_proj = 160, 40
    det = capi.DetectorGeometry(n, 144, 0.4, 0.4, 6.0, 0.0, 500.0, 500.0, 360.0 / n_proj)
    vol = capi.calculate_volume_geometry(det)
    roi = capi.Roi(21, 101, 30, 130, 7, 7 + 97)
    reg = capi.apply_roi(vol, roi)
    region = (reg.dim_x, reg.dim_y, reg.dim_z)
    want = _one_piece(ctx, det, vol, n_proj, roi=roi, region=region)
    got, _ = _run_in_process(3, det, vol, n_proj, roi=roi, first_round=6, max_round=12)
    assert np.isfinite(got).all() and np.array_equal(got, want)


@pytest.mark.parametrize("e2e", [True, False])
def test_members_own_boxes_in_x_and_z(ctx, e2e):
    """x_parts = 2 with four members: two z-runs of twice the height, each cut in two along x (what keeps the
    128-slice tiles on volumes with few slices per GPU); boxes land inside one region-wide host volume"""
    det, vol, n_proj = _case(k=64, extra_z=3)
    want = _one_piece(ctx, det, vol, n_proj)
    got, _ = _run_in_process(4, det, vol, n_proj, e2e=e2e, x_parts=2, first_round=8, max_round=16)
    assert np.isfinite(got).all() and np.array_equal(got, want)
    # ROI with odd offsets on top of the x-parts, three x-parts with a remainder
    roi = capi.Roi(5, 60, 3, 50, 2, 2 + 40)
    reg = capi.apply_roi(vol, roi)
    want = _one_piece(ctx, det, vol, n_proj, roi=roi, region=(reg.dim_x, reg.dim_y, reg.dim_z))
    got, _ = _run_in_process(3, det, vol, n_proj, roi=roi, e2e=e2e, x_parts=3, first_round=6, max_round=12)
    assert np.isfinite(got).all() and np.array_equal(got, want)


def test_whole_projection_exchange_is_the_same_volume(ctx):
    det, vol, n_proj = _case()
    want = _one_piece(ctx, det, vol, n_proj)
    got, stats = _run_in_process(2, det, vol, n_proj, whole_projections=True, first_round=8, max_round=16)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("world,e2e", [(1, True), (1, False), (2, True)])
def test_sixteen_bit_samples_equal_the_same_counts_as_floats(ctx, world, e2e):
    """Detector-native 16-bit counts (PARIS_B200_SAMPLES_U16: half the upload, widened by the filter kernel's first
    load) against the very same counts handed over as floats (what the reference's reader produces on the host,
    src/his.cpp:169-185): bit for bit, from host pointers and from device-resident samples."""
    det, vol, n_proj = _case()
    scale = phantom.counts_scale(_ellipsoids(det))          # the longest chord of the phantom stays inside 16 bits
    as_f32, _ = _run_in_process(world, det, vol, n_proj, e2e=e2e, counts_scale=scale, first_round=8, max_round=16)
    as_u16, _ = _run_in_process(world, det, vol, n_proj, e2e=e2e, counts_scale=scale, first_round=8, max_round=16,
                                sample_type=capi.SAMPLES_U16)
    assert np.isfinite(as_f32).all() and np.abs(as_f32).max() > 10.0
    assert np.array_equal(as_u16, as_f32)


def test_sixteen_bit_stage_call_equals_float_stage_call(ctx):
    """paris_b200_filter_to_stack_batch_u16 against paris_b200_filter_to_stack_batch on the same counts, both stack
    layouts, a batch of more projections than one launch takes"""
    rng = np.random.default_rng(5)
    n_proj = 70
    for n, layout in ((128, capi.LAYOUT_PLAIN), (320, capi.LAYOUT_SPLIT2)):
        det = capi.DetectorGeometry(n, n - 6, 0.4, 0.4, 0.0, 0.0, 500.0, 500.0, 360.0 / n_proj)
        px = det.n_row * det.n_col
        counts = rng.integers(0, 65536, size=(n_proj, det.n_col, det.n_row), dtype=np.uint16)
        counts[0, 0, :4] = (0, 1, 65535, 32768)
        as_float = counts.astype(np.float32)
        d_u16, d_f32 = ctx.dev_alloc(n_proj * px * 2), ctx.dev_alloc(n_proj * px * 4)
        ctx.vol_h2d(counts.view(np.float32), d_u16, counts.size // 2)
        ctx.vol_h2d(as_float, d_f32, as_float.size)
        filt = ctx.filter_create(capi.filter_size(det.n_row), float(det.l_px_row))
        slot_bytes, _ = capi.stack_slot_bytes(det.n_row, det.n_col)
        stacks = []
        for u16 in (False, True):
            stack = ctx.stack_alloc(det.n_row, det.n_col, n_proj)
            ctx.vol_h2d(np.zeros(n_proj * slot_bytes // 4, np.float32), stack, n_proj * slot_bytes // 4)  # the pad columns
            if u16:
                ctx.filter_to_stack_batch_u16(d_u16, px, n_proj, det, filt, stack, 0, layout)
            else:
                ctx.filter_to_stack_batch(d_f32, px, n_proj, det, filt, stack, 0, layout)
            out = np.empty(n_proj * slot_bytes // 4, np.float32)
            ctx.vol_d2h(stack, out, out.size)
            stacks.append(out)
            ctx.stack_free(stack)
        assert np.abs(stacks[0]).max() > 0 and np.array_equal(stacks[0], stacks[1])
        ctx.filter_destroy(filt)
        ctx.dev_free(d_u16)
        ctx.dev_free(d_f32)


# ---- one process per GPU over CUDA IPC (needs two devices) ---------------------------------------------------------

def _ipc_worker(rank, world, conns, out_dir, exchange):
    import numpy as np
    from paris_b200 import capi
    from paris_b200.multi import GroupMember
    det, vol, n_proj = _case()
    m = GroupMember(rank, rank, world, det, vol, n_proj, first_round=8, max_round=16, exchange=exchange)
    parent = conns[rank]
    parent.send(m.export())
    m.connect(parent.recv())
    m.generate_inputs(_ellipsoids(det))
    _poison(m)
    parent.send("ready")
    assert parent.recv() == "go"
    m.alloc_host_slabs()
    for _ in range(2):
        m.step_e2e()
    np.save(os.path.join(out_dir, f"slab{rank}.npy"), m.h_slabs.array)
    parent.send((m.info.z_first, m.info.z_count, m.group.info().bytes_pushed, m.group.info().memops))
    assert parent.recv() == "close"      # nobody tears its stack down while a peer may still push into it
    m.close()


@pytest.mark.parametrize("exchange", [capi.EXCHANGE_COPY_ENGINE, capi.EXCHANGE_KERNEL])
def test_two_processes_two_gpus_over_ipc(ctx, tmp_path, exchange):
    if capi.device_count() < 2:
        pytest.skip("needs two GPUs")
    world = 2
    det, vol, n_proj = _case()
    want = _one_piece(ctx, det, vol, n_proj)
    spawn = mp.get_context("spawn")
    pipes = [spawn.Pipe() for _ in range(world)]
    procs = [spawn.Process(target=_ipc_worker, args=(r, world, [p[1] for p in pipes], str(tmp_path), exchange))
             for r in range(world)]
    for p in procs:
        p.start()
    handles = [pipes[r][0].recv() for r in range(world)]
    for r in range(world):
        pipes[r][0].send(handles)
    for r in range(world):
        assert pipes[r][0].recv() == "ready"
    for r in range(world):
        pipes[r][0].send("go")
    got = np.full((vol.dim_z, vol.dim_y, vol.dim_x), np.nan, np.float32)
    for r in range(world):
        z_first, z_count, pushed, memops = pipes[r][0].recv()
        got[z_first:z_first + z_count] = np.load(os.path.join(str(tmp_path), f"slab{r}.npy"))
        print(f"rank {r}: slices [{z_first}, {z_first + z_count}), pushed {pushed / 1e6:.1f} MB, memops {memops}")
    for r in range(world):
        pipes[r][0].send("close")
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert np.isfinite(got).all() and np.array_equal(got, want)
