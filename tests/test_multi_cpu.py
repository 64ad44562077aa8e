"""The N > 1 decomposition on CPU.

The group's plan -- who filters which projections, who owns which slices, which detector rows travel to whom -- is
host arithmetic behind the C ABI (paris_b200_group_plan / _share, csrc/group.cu) and needs no GPU.  Here it is checked
(a) for its invariants on BASELINE's configurations, (b) against the reference's slab arithmetic, and (c) end to end
over gloo with world_size 2 and 3: every rank filters ITS share with the CPU oracle, sends every peer only the BAND of
detector rows the plan says that peer needs (everything else stays NaN), backprojects all projections into its slabs
through the oracle, and the slabs reassembled by offset equal the one-piece reconstruction bit for bit -- so a band
that is too small, a projection filtered twice or not at all, or a slab at the wrong offset cannot go unnoticed."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from paris_b200 import capi
from paris_b200.multi import SlabPlan

from cases import both_det, coarse_volume, shepp_logan, to_capi_vol

N_PROJ = 14


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _case(P):
    odet, det = both_det(48, 40, n_proj=N_PROJ, delta_s=1.0)
    ovol = coarse_volume(odet, 24)
    return odet, det, ovol, to_capi_vol(ovol)


def _worker(rank, world, port_no, out_path, spm):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        P = oracle.Port()
        odet, det, ovol, vol = _case(P)
        cfg = capi.group_config(rank, world, det, vol, N_PROJ, slabs_per_member=spm, first_round=4, max_round=6)
        plan = capi.group_plan(cfg)
        raw = shepp_logan(odet, N_PROJ)
        # my stack: NaN wherever nothing was put
        stack = np.full((N_PROJ, odet.n_col, odet.n_row), np.nan, np.float32)
        for rd in range(plan.rounds):
            first, count = capi.group_share(plan, world, rd, rank)
            mine = {i: P.filter(P.weight(raw[i], odet), odet) for i in range(first, first + count)}
            for i, f in mine.items():
                stack[i] = f
            # the exchange of this round: every peer gets the rows of ITS band, nothing else
            outbox = [None] * world
            for k in range(world):
                lo, hi = plan.band_lo[k], min(plan.band_hi[k], odet.n_col)
                outbox[k] = {i: f[lo:hi].copy() for i, f in mine.items()} if k != rank else {}
            # (gloo has no object all-to-all: gather every outbox, keep what is addressed to me)
            everything = [None] * world
            dist.all_gather_object(everything, outbox)
            lo, hi = plan.band_lo[rank], min(plan.band_hi[rank], odet.n_col)
            for src in range(world):
                if src != rank:
                    for i, rows in everything[src][rank].items():
                        stack[i, lo:hi] = rows
        # every member backprojects ALL projections into each of its slabs
        total = plan.slabs_total
        parts = []
        for s in range(spm):
            sid = rank * spm + s
            dz = plan.slab_dz + (plan.slab_remainder if sid == total - 1 else 0)
            off = sid * plan.slab_dz
            slab = np.zeros((dz, ovol.dim_y, ovol.dim_x), np.float32)
            for i in range(N_PROJ):
                P.backproject(stack[i], i, slab, odet, ovol, v_offset=off)
            parts.append((off, slab))
        gathered = [None] * world
        dist.gather_object(parts, gathered if rank == 0 else None, dst=0)
        if rank == 0:
            out = np.full((ovol.dim_z, ovol.dim_y, ovol.dim_x), np.nan, np.float32)
            for member in gathered:
                for off, slab in member:
                    out[off:off + slab.shape[0]] = slab
            np.save(out_path, out)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,spm", [(2, 1), (3, 1), (2, 3)])
def test_band_exchange_over_gloo_equals_one_piece(tmp_path, world, spm, port):
    out = str(tmp_path / "vol.npy")
    mp.spawn(_worker, args=(world, _free_port(), out, spm), nprocs=world, join=True)
    got = np.load(out)
    odet, det, ovol, vol = _case(port)
    ref, _ = port.reconstruct(shepp_logan(odet, N_PROJ), (ovol.dim_z, ovol.dim_y, ovol.dim_x), odet, ovol)
    assert np.isfinite(got).all(), "a voxel read a detector row outside the band its member received"
    assert np.array_equal(got, ref)


def _baseline(cfg_name):
    import bench
    return bench.geometry(cfg_name)


@pytest.mark.parametrize("name,world,spm", [("c2", 8, 1), ("c3", 8, 1), ("c3", 2, 1), ("c4", 8, 1), ("c5", 8, 2), ("c5", 4, 4),
                                            ("c1", 3, 1)])
def test_plan_invariants_on_baseline_configurations(port, name, world, spm):
    det, vol, n_proj, roi, region = _baseline(name)
    plans = [capi.group_plan(capi.group_config(r, world, det, vol, n_proj, roi=roi, slabs_per_member=spm))
             for r in range(world)]
    p = plans[0]
    for q in plans[1:]:       # every member computes the same plan
        assert bytes(q) == bytes(p)
    assert (p.region_x, p.region_y, p.region_z) == tuple(region)
    # slabs: the reference's arithmetic (src/cuda/subvolume_information.cpp:112-116, src/make_volume.cpp:32-34)
    info = port.make_subvolume_information(oracle.VolumeGeometry(p.region_x, p.region_y, p.region_z, 1, 1, 1), world * spm)
    assert (p.slabs_total, p.slab_dz, p.slab_remainder) == (world * spm, info.dim_z, info.remainder)
    # rounds: consecutive, cover the scan once, shares partition every round in rank order
    seen = []
    for rd in range(p.rounds):
        assert p.round_first[rd] == len(seen)
        at = p.round_first[rd]
        for r in range(world):
            first, count = capi.group_share(p, world, rd, r)
            assert first == at
            at += count
        assert at == p.round_first[rd] + p.round_count[rd]
        seen += list(range(p.round_first[rd], at))
    assert seen == list(range(n_proj))
    assert p.round_count[0] <= 64 or p.rounds == 1
    # bands: multiples of 8 rows inside the line, and together they are far less than world whole projections
    total_rows = 0
    for r in range(world):
        lo, hi = p.band_lo[r], p.band_hi[r]
        assert lo % 8 == 0 and hi % 8 == 0 and lo < hi <= p.pitch
        total_rows += hi - lo
    if world >= 4:
        assert total_rows < 0.45 * world * det.n_col, "the exchange should move a fraction of an all-gather"


@pytest.mark.parametrize("name,world", [("c3", 8), ("c4", 8), ("c5", 8), ("c2", 4)])
def test_bands_cover_every_row_a_slab_can_read(name, world):
    """brute force in float64: rows floor(v), floor(v) + 1 of every slice of a member's slabs at the extreme
    magnifications of the region lie inside the member's band"""
    det, vol, n_proj, roi, region = _baseline(name)
    p = capi.group_plan(capi.group_config(0, world, det, vol, n_proj, roi=roi))
    x1 = roi.x1 if roi is not None else 0
    y1 = roi.y1 if roi is not None else 0

    def centred(i, dim, size):
        return -(dim * size / 2.0) + size / 2.0 + i * size
    xs = [centred(i, vol.dim_x, vol.l_vx_x) for i in (x1, x1 + region[0] - 1)]
    ys = [centred(i, vol.dim_y, vol.l_vx_y) for i in (y1, y1 + region[1] - 1)]
    r = max(np.hypot(x, y) for x in xs for y in ys)
    d_sd = abs(det.d_so) + abs(det.d_od)
    fs = np.array([d_sd / (det.d_so + r), d_sd / (det.d_so - r)])
    min_v = -(det.n_col * det.l_px_col / 2.0) - det.delta_t * det.l_px_col
    for k in range(world):
        z0 = p.region_z0 + k * p.slab_dz
        dz = p.slab_dz + (p.slab_remainder if k == world - 1 else 0)
        z = centred(np.arange(z0, z0 + dz), vol.dim_z, vol.l_vx_z)
        v = (z[:, None] * fs[None, :] - min_v) / det.l_px_col - 0.5
        rows = np.floor(v)
        used = (rows >= 0) & (rows + 1 < det.n_col)      # the reference reads nothing otherwise (:65-71)
        if used.any():
            assert rows[used].min() >= p.band_lo[k] and rows[used].max() + 1 < p.band_hi[k]


@pytest.mark.parametrize("dim_z,world", [(512, 8), (514, 8), (50, 3), (7, 7), (1029, 4)])
def test_slab_plan_matches_reference_split(port, dim_z, world):
    """SlabPlan == make_subvolume_information + make_volume(last) + offset = id * dim_z of the reference."""
    info = port.make_subvolume_information(oracle.VolumeGeometry(8, 8, dim_z, 1, 1, 1), world)
    covered = 0
    for r in range(world):
        p = SlabPlan(dim_z, world, r)
        assert p.offset == r * info.dim_z
        assert p.slab_dz == info.dim_z + (info.remainder if r == world - 1 else 0)
        assert p.offset == covered
        covered += p.slab_dz
    assert covered == dim_z
    with pytest.raises(ValueError):
        SlabPlan(3, 4, 0)


@pytest.mark.parametrize("name,world,xp", [("c2", 8, 2), ("c1", 8, 8), ("c3", 8, 1), ("c2", 4, 4)])
def test_plan_with_x_parts(name, world, xp):
    """x_parts cuts the region along x as well: world / x_parts z-runs, members of one z-run share its band"""
    det, vol, n_proj, roi, region = _baseline(name)
    p = capi.group_plan(capi.group_config(0, world, det, vol, n_proj, roi=roi, x_parts=xp))
    assert (p.x_parts, p.x_dx * xp + p.x_remainder) == (xp, region[0])
    assert p.slabs_total == world // xp and p.slab_dz * p.slabs_total + p.slab_remainder == region[2]
    for k in range(world):
        assert (p.band_lo[k], p.band_hi[k]) == (p.band_lo[(k // xp) * xp], p.band_hi[(k // xp) * xp])
    with pytest.raises(capi.Error):
        capi.group_plan(capi.group_config(0, world, det, vol, n_proj, roi=roi, x_parts=3))


def test_plan_rejects_more_slabs_than_slices(port):
    odet, det = both_det(32, 8, n_proj=8)
    vol = to_capi_vol(port.calculate_volume_geometry(odet))
    with pytest.raises(capi.Error):
        capi.group_plan(capi.group_config(0, vol.dim_z + 1, det, vol, 8))
