"""The command-line driver end to end on the GPU: HIS files in, DDBVF out, checked against the
CPU oracle on the same projections; the projection source against the reference's own source."""
from __future__ import annotations

import os
import subprocess

import numpy as np
import pytest

import oracle
from oracle import formats
from paris_b200 import io as pio
import cases

pytestmark = pytest.mark.gpu

N_ROW, N_COL, N_PROJ = 64, 48, 40


def write_geometry(path, det):
    with open(path, "w") as f:
        f.write("# test geometry\n")
        for k in ("n_row", "n_col", "l_px_row", "l_px_col", "delta_s", "delta_t", "d_so", "d_od", "delta_phi"):
            f.write(f"{k} = {getattr(det, k)!r}\n")


def make_scan(tmp_path, number_type=128, delta_s=0.0, frames_per_file=(7, 13, 20)):
    odet, _ = cases.both_det(N_ROW, N_COL, 0.4, delta_s=delta_s, n_proj=N_PROJ)
    stack = cases.shepp_logan(odet, N_PROJ)                       # (n_proj, n_col, n_row) float32
    if number_type == 4:
        stack = np.round(stack * (60000.0 / stack.max())).astype(np.float32)   # what a 16-bit detector delivers
    d = tmp_path / "scan"
    d.mkdir()
    first = 0
    for i, n in enumerate(frames_per_file):
        formats.write_his(str(d / f"proj_{i:03d}.his"), stack[first:first + n], number_type)
        first += n
    assert first == N_PROJ
    (d / "proj_001b_broken.his").write_bytes(b"not a his file at all")     # skipped with a warning (src/source.cpp:97)
    geo = tmp_path / "geometry.cfg"
    write_geometry(str(geo), odet)
    return odet, stack, str(d), str(geo)


def run_cli(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([pio.CLI_PATH] + args, capture_output=True, text=True, env=e, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    return r


def oracle_volume(port, odet, stack, indices=None, angles=None, roi=None):
    full = port.calculate_volume_geometry(odet)
    region = full if roi is None else port.apply_roi(full, roi)
    vol = np.zeros((region.dim_z, region.dim_y, region.dim_x), np.float32)
    idx = range(stack.shape[0]) if indices is None else indices
    for i in idx:
        p = port.filter(port.weight(stack[i], odet), odet)
        port.backproject(p, i, vol, odet, full, roi=roi, phi_deg=None if angles is None else float(angles[i]))
    return vol


def check(got, want, n_used):
    c = cases.contrast(n_used) * (want.max() - want.min() > 0)
    # contrast in the units of this input: scale by the data (u16 scans are not in phantom units)
    scale = max(float(np.abs(want).max()), 1e-30)
    mx = float(np.abs(got.astype(np.float64) - want).max()) / scale
    rms = float(np.sqrt(np.mean((got.astype(np.float64) - want) ** 2))) / scale
    assert mx <= cases.MAX_ABS_TOL and rms <= cases.RMSE_TOL, (mx, rms, c)


def test_cli_reconstructs_his_scan_to_ddbvf(tmp_path, port):
    odet, stack, scan, geo = make_scan(tmp_path)
    out = tmp_path / "out"
    r = run_cli(["--geometry", geo, "--input", scan, "--output", str(out)])
    n_dev = pio.capi.device_count()
    assert (f"Created {n_dev} tasks for {n_dev} devices" if n_dev > 1 else "Created 1 task for 1 device") in r.stderr
    got = formats.read_ddbvf(str(out / "vol.ddbvf"))
    want = oracle_volume(port, odet, stack)
    assert got.shape == want.shape
    check(got, want, N_PROJ)


def test_cli_u16_scan_roi_and_name(tmp_path, port):
    odet, stack, scan, geo = make_scan(tmp_path, number_type=4, delta_s=2.0)
    roi = oracle.Roi(5, 40, 8, 50, 3, 30)
    out = tmp_path / "out"
    run_cli(["--geometry", geo, "--input", scan, "--output", str(out), "--name", "roi_scan", "--roi",
             "--roi-x1", "5", "--roi-x2", "40", "--roi-y1", "8", "--roi-y2", "50", "--roi-z1", "3", "--roi-z2", "30"])
    got = formats.read_ddbvf(str(out / "roi_scan.ddbvf"))
    want = oracle_volume(port, odet, stack, roi=roi)
    assert got.shape == want.shape == (27, 42, 35)
    check(got, want, N_PROJ)


def test_cli_angle_file_and_quality(tmp_path, port):
    odet, stack, scan, geo = make_scan(tmp_path)
    rng = np.random.default_rng(5)
    angles = (np.arange(N_PROJ) * odet.delta_phi + rng.uniform(-1.0, 1.0, N_PROJ)).astype(np.float32)
    af = tmp_path / "angles.txt"
    af.write_text("\n".join(repr(float(a)) for a in angles) + "\n")
    out = tmp_path / "out"
    run_cli(["--geometry", geo, "--input", scan, "--output", str(out), "--angles", str(af), "--quality", "3"])
    got = formats.read_ddbvf(str(out / "vol.ddbvf"))
    used = list(range(0, N_PROJ, 3))                               # src/source.cpp:105
    want = oracle_volume(port, odet, stack, indices=used, angles=angles)
    check(got, want, len(used))


def test_cli_forced_slabs_are_bit_identical(tmp_path):
    _, _, scan, geo = make_scan(tmp_path)
    run_cli(["--geometry", geo, "--input", scan, "--output", str(tmp_path / "one")])
    r = run_cli(["--geometry", geo, "--input", scan, "--output", str(tmp_path / "three")], env={"PARIS_B200_SLABS": "3"})
    assert "Created 3 tasks for" in r.stderr
    a = formats.read_ddbvf(str(tmp_path / "one" / "vol.ddbvf"))
    b = formats.read_ddbvf(str(tmp_path / "three" / "vol.ddbvf"))
    assert np.array_equal(a, b)


@pytest.mark.parametrize("quality", [1, 2, 3])
def test_source_walk_equals_reference(tmp_path, quality):
    _, stack, scan, _ = make_scan(tmp_path)
    idx, phi, from_file, first = pio.source_walk(scan, None, quality)
    assert list(idx) == list(range(0, N_PROJ, quality))
    assert np.array_equal(first, stack[idx, 0, 0]) and not from_file.any() and not phi.any()
    if oracle.have_ref():
        ridx, rphi, rfirst = formats.RefIO().source_walk(scan, None, quality)
        assert np.array_equal(idx, ridx) and np.array_equal(first, rfirst) and np.array_equal(phi, rphi)


def test_source_angles_and_second_walk_restarts_counting(tmp_path):
    _, _, scan, _ = make_scan(tmp_path)
    af = tmp_path / "angles.txt"
    af.write_text(" ".join(str(0.5 * i) for i in range(30)) + "\n")     # shorter than the scan
    idx, phi, from_file, _ = pio.source_walk(scan, str(af), 1)
    assert list(idx) == list(range(N_PROJ))
    assert np.array_equal(phi[:30], np.float32([0.5 * i for i in range(30)])) and from_file[:30].all()
    assert not from_file[30:].any()                                   # callers fall back to idx*delta_phi there
    idx2, _, _, _ = pio.source_walk(scan, None, 1)                    # (the reference would continue at N_PROJ: F9)
    assert list(idx2) == list(range(N_PROJ))


# ---- the group driver (several members; on a one-GPU box they share the device) -------------------------------------

@pytest.mark.parametrize("members,slabs", [(1, 1), (3, 3), (2, 4)])
def test_cli_group_equals_the_task_model_bit_for_bit(tmp_path, members, slabs):
    """PARIS_B200_GROUP: every member reads and filters 1/N of the frames, detector-row bands travel between the
    members, every member backprojects everything into its own slabs -- the same file as the reference's task model
    (every device reads and filters everything) writes."""
    _, _, scan, geo = make_scan(tmp_path)
    run_cli(["--geometry", geo, "--input", scan, "--output", str(tmp_path / "tasks")], env={"PARIS_B200_TASKS": "1"})
    r = run_cli(["--geometry", geo, "--input", scan, "--output", str(tmp_path / "group")],
                env={"PARIS_B200_GROUP": "1", "PARIS_B200_GROUP_MEMBERS": str(members), "PARIS_B200_SLABS": str(slabs),
                     "PARIS_B200_GROUP_TIMEOUT_S": "20"})
    assert f"Group of {members} member" in r.stderr
    assert r.stderr.count("projections filtered here") == members
    a = formats.read_ddbvf(str(tmp_path / "tasks" / "vol.ddbvf"))
    b = formats.read_ddbvf(str(tmp_path / "group" / "vol.ddbvf"))
    assert np.array_equal(a, b)


def test_cli_group_u16_roi_angles_quality_against_the_oracle(tmp_path, port):
    odet, stack, scan, geo = make_scan(tmp_path, number_type=4, delta_s=2.0)
    rng = np.random.default_rng(6)
    angles = (np.arange(N_PROJ) * odet.delta_phi + rng.uniform(-1.0, 1.0, N_PROJ)).astype(np.float32)
    af = tmp_path / "angles.txt"
    af.write_text("\n".join(repr(float(a)) for a in angles[:31]) + "\n")     # shorter than the scan: idx*delta_phi after it
    roi = oracle.Roi(5, 40, 8, 50, 3, 30)
    out = tmp_path / "out"
    args = ["--geometry", geo, "--input", scan, "--angles", str(af), "--quality", "2", "--roi",
            "--roi-x1", "5", "--roi-x2", "40", "--roi-y1", "8", "--roi-y2", "50", "--roi-z1", "3", "--roi-z2", "30"]
    env = {"PARIS_B200_GROUP": "1", "PARIS_B200_GROUP_MEMBERS": "3", "PARIS_B200_GROUP_TIMEOUT_S": "20"}
    r = run_cli(args + ["--output", str(out)], env=env)
    # a scan of 16-bit files is uploaded as it is and widened by the filter kernel ...
    assert r.stderr.count("16-bit projections filtered here") == 3
    got = formats.read_ddbvf(str(out / "vol.ddbvf"))
    # ... which is the same volume, bit for bit, as widening on the host like src/his.cpp:98-99
    r = run_cli(args + ["--output", str(tmp_path / "widened")], env=dict(env, PARIS_B200_SAMPLES="f32"))
    assert "16-bit projections" not in r.stderr
    assert np.array_equal(got, formats.read_ddbvf(str(tmp_path / "widened" / "vol.ddbvf")))
    used = list(range(0, N_PROJ, 2))
    full_angles = np.array([angles[i] if i < 31 else np.float32(i) * np.float32(odet.delta_phi) for i in range(N_PROJ)],
                           np.float32)
    want = oracle_volume(port, odet, stack, indices=used, angles=full_angles, roi=roi)
    assert got.shape == want.shape
    check(got, want, len(used))
