"""The N > 1 host logic on CPU: world_size 2 and 3 over gloo.  Every rank filters its block of projections,
the filtered stack is all-gathered exactly the way paris_b200.multi does it (equal chunks, in place), every
rank backprojects ALL projections into its own z-slab, and the slabs reassembled by offset equal the
one-piece reconstruction bit for bit.  The arithmetic is the CPU oracle's (there is no GPU here); the
partitioning, chunking and offsets are the product's (paris_b200.multi.SlabPlan)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from paris_b200.multi import SlabPlan

from cases import both_det, shepp_logan

N_PROJ = 10


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port_no, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        P = oracle.Port()
        odet, _ = both_det(40, 36, n_proj=N_PROJ)
        vg = P.calculate_volume_geometry(odet)
        plan = SlabPlan(vg.dim_z, world, rank)
        lo, hi, chunk = plan.projection_block(N_PROJ)
        stack = shepp_logan(odet, N_PROJ)
        px = odet.n_row * odet.n_col
        # the stack has world*chunk slots; slots beyond N_PROJ stay zero and are never backprojected
        full = torch.zeros(world * chunk * px, dtype=torch.float32)
        for i in range(lo, hi):
            full[i * px:(i + 1) * px] = torch.from_numpy(P.filter(P.weight(stack[i], odet), odet).ravel())
        mine = full[rank * chunk * px:(rank + 1) * chunk * px]
        dist.all_gather_into_tensor(full, mine.clone())
        filtered = full.numpy().reshape(world * chunk, odet.n_col, odet.n_row)
        slab = np.zeros((plan.slab_dz, vg.dim_y, vg.dim_x), np.float32)
        for i in range(N_PROJ):
            P.backproject(filtered[i], i, slab, odet, vg, v_offset=plan.offset)
        # host reassembly by offset: gather the slabs on rank 0
        slabs = [None] * world
        dist.gather_object((plan.offset, slab), slabs if rank == 0 else None, dst=0)
        if rank == 0:
            vol = np.zeros((vg.dim_z, vg.dim_y, vg.dim_x), np.float32)
            for off, s in slabs:
                vol[off:off + s.shape[0]] = s
            np.save(out_path, vol)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_slab_allgather_pipeline_over_gloo(tmp_path, world, port):
    out = str(tmp_path / "vol.npy")
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    got = np.load(out)
    odet, _ = both_det(40, 36, n_proj=N_PROJ)
    vg = port.calculate_volume_geometry(odet)
    stack = shepp_logan(odet, N_PROJ)
    ref, _ = port.reconstruct(stack, (vg.dim_z, vg.dim_y, vg.dim_x), odet, vg)
    assert np.array_equal(got, ref)


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


@pytest.mark.parametrize("n_proj,world", [(720, 8), (1440, 8), (10, 3), (5, 8)])
def test_projection_blocks_partition_the_scan(n_proj, world):
    seen = []
    for r in range(world):
        lo, hi, chunk = SlabPlan(64, world, r).projection_block(n_proj)
        assert hi - lo <= chunk and chunk * world >= n_proj
        seen += list(range(lo, hi))
    assert seen == list(range(n_proj))


@pytest.mark.parametrize("n_proj,world", [(720, 8), (720, 4), (720, 2), (1440, 8), (2880, 8), (64, 2), (12, 3)])
def test_block_cyclic_rounds_keep_projection_order(n_proj, world):
    """Pipelined exchange: every round is world*m CONSECUTIVE projections, rank r owns the r-th block of m;
    all ranks together cover the scan exactly once and a round fits one backprojection launch."""
    m = SlabPlan(64, world, 0).cyclic_blocks(n_proj, 64)
    assert m >= 1 and world * m <= 64 and (n_proj // world) % m == 0
    rounds = n_proj // (world * m)
    owner = {}
    for r in range(world):
        for local in range(n_proj // world):
            rd, j = divmod(local, m)
            g = rd * world * m + r * m + j
            assert g not in owner
            owner[g] = r
    assert sorted(owner) == list(range(n_proj))
    for rd in range(rounds):
        block = [owner[g] for g in range(rd * world * m, (rd + 1) * world * m)]
        assert block == [r for r in range(world) for _ in range(m)]


def test_block_cyclic_falls_back_when_uneven():
    assert SlabPlan(64, 8, 0).cyclic_blocks(721, 64) == 0


@pytest.mark.parametrize("n_proj,world,first,large", [(720, 8, 64, 256), (720, 2, 64, 256), (1440, 8, 64, 256),
                                                      (2880, 8, 64, 256), (48, 2, 8, 16), (12, 3, 64, 256), (720, 4, 64, 64)])
def test_round_schedule_covers_the_scan_in_order(n_proj, world, first, large):
    """Growing rounds: round r is world*m_r CONSECUTIVE projections, rank k owns the k-th block of m_r; together the
    rounds cover the scan exactly once, in order; no round exceeds one batch (plus the merged sliver)."""
    from paris_b200.multi import round_schedule
    ms = round_schedule(n_proj, world, first, large)
    assert sum(ms) * world == n_proj and all(m >= 1 for m in ms)
    assert ms[0] * world <= max(first, world) or len(ms) == 1
    assert all(m * world <= large * 1.25 + world for m in ms)
    assert all(b >= a for a, b in zip(ms[:-2], ms[1:-1]))          # non-decreasing up to the last (remainder) round
    start, seen = 0, []
    for m in ms:
        for k in range(world):
            seen += list(range(start + k * m, start + (k + 1) * m))
        start += world * m
    assert seen == list(range(n_proj))
    assert round_schedule(721, 8) == []
