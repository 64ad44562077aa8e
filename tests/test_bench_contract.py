"""The JSON line bench.py prints must keep the driver's contract.  Checked on the committed lines under profiles/
(produced by the real runs on B200 boxes), so a change to bench.py that drops or renames a key is caught on CPU."""
import glob
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LINES = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_bench_*.json")) + glob.glob(os.path.join(ROOT, "profiles", "r*_scale_*.json")))


def last_line(path):
    return [json.loads(l) for l in open(path) if l.startswith("{")][-1]


def test_profiles_hold_bench_lines():
    assert LINES, "profiles/ holds no bench line"


@pytest.mark.parametrize("path", LINES, ids=[os.path.basename(p) for p in LINES])
def test_bench_line_keeps_the_contract(path):
    d = last_line(path)
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "e2e", "gpu_launches", "clocks"):
        assert key in d, key
    assert d["unit"] == "GUPS" and d["higher_is_better"] is True and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert d["vs_baseline"] is None                      # BASELINE.md holds no published number
    assert "workload" in d["config"] and "model" not in d["config"]
    r = d["roofline"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in r, key
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    e = d["e2e"]
    for key in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert key in e, key
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"] * 1.001
    assert d["gpu_launches"] > 0
    assert d["steps"] >= 1 and d["warmup"] >= 3
    # whole-job throughput = updates / time
    assert d["value"] > 0 and d["ms_per_step"] > 0
    if d["n_gpus"] == 1 and "cpu_baseline" in d:
        c = d["cpu_baseline"]
        for key in ("value", "unit", "cores", "kind", "sample"):
            assert key in c, key
        assert c["kind"] in ("reference", "port") and c["cores"] >= 1


def test_final_single_gpu_line_has_clocks_and_cpu_baseline():
    d = last_line(os.path.join(ROOT, "profiles", "r1_bench_c2_1gpu.json"))
    assert d["clocks"]["sm_mhz"] and d["clocks"]["sm_max_mhz"] and isinstance(d["clocks"]["reasons"], list)
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["value"] > 0
    assert "512^3" in d["config"]["workload"]


def test_round_two_lines_are_on_the_north_star_config_with_green_parity():
    d = last_line(os.path.join(ROOT, "profiles", "r2_bench_c3_1gpu.json"))
    assert "1024^3" in d["config"]["workload"] and "2048^2" in d["config"]["workload"]      # BASELINE.json's configs[2]
    assert d["roofline"]["kernel"].startswith("bp_tma_kernel<") and d["roofline"]["launches_by_kernel"]["exact"] == 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(d["clocks"]["reasons"])
    many = [p for p in LINES if os.path.basename(p).startswith("r2_scale_")]
    assert many
    for path in many:
        m = last_line(path)
        assert m["n_gpus"] > 1 and m["scaling"] == "strong"
        # every multi-GPU line certifies its volume: bands of every rank's slab against the exact kernel
        assert m["parity_check"]["ok"] is True and m["parity_check"]["max"] <= 1e-4 and m["parity_check"]["rmse"] <= 1e-5
        assert m["exchange"]["bytes_pushed_per_step_rank0"] < m["exchange"]["all_gather_bytes_per_step_rank0"]


def test_boxes_of_the_gpus_are_balanced_from_four_gpus_on():
    """bench.auto_x_parts: z-slabs up to two GPUs (two halves are mirror images), two mirror z-runs x N/2 x-parts from
    four GPUs on for full regions, z-slabs for ROI regions and streamed slabs, and never fewer than 128 slices per
    GPU where the GPU count allows it; the plan accepts every choice."""
    import bench
    from paris_b200 import capi
    assert [bench.auto_x_parts(n, (1024, 1024, 1024), 1, True) for n in (1, 2, 4, 8)] == [1, 1, 2, 4]
    assert [bench.auto_x_parts(n, (512, 512, 512), 1, True) for n in (1, 2, 4, 8)] == [1, 1, 2, 4]
    assert bench.auto_x_parts(8, (1024, 1024, 1024), 1, False) == 1          # ROI regions keep their z-slabs
    assert bench.auto_x_parts(8, (2048, 2048, 2048), 2, True) == 1           # streamed slabs (config 5)
    assert bench.auto_x_parts(8, (512, 512, 512), 1, False) == 2             # 64 slices per GPU otherwise
    assert bench.auto_x_parts(3, (1024, 1024, 1024), 1, True) == 1
    for name in ("c2", "c3"):
        det, vol, n_proj, roi, dims = bench.geometry(name)
        for world in (4, 8):
            xp = bench.auto_x_parts(world, dims, 1, roi is None)
            plans = [capi.group_plan(capi.group_config(r, world, det, vol, n_proj, roi=roi, x_parts=xp)) for r in range(world)]
            assert all(p.x_parts == xp and p.slabs_total == world // xp for p in plans)
            assert plans[0].slab_dz * (world // xp) == dims[2] and plans[0].x_dx * xp == dims[0]


def test_reference_arm_runs_here_and_loads_nothing_of_the_product():
    """`bench.py --impl reference` is CPU work (the reference's OpenMP backend from oracle/_ref, else the oracle port): it
    runs in this container, prints the contract's line with the reference-arm keys, and -- asserted inside bench.py --
    never imports paris_b200."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--config", "c1", "--steps", "2",
                        "--warmup", "1", "--cpu-budget", "0.5"], capture_output=True, text=True, timeout=300, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["metric"] == "fdk_reconstruction_gups" and line["unit"] == "GUPS"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["steps"] == 2 and line["warmup"] == 1
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"] == line["e2e"]["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
