"""Turn the ncu artefacts in gpurun_out/ into the tracked summaries under profiles/ (run where ncu is installed).

    python scripts/make_profile_summary.py r1
"""
import collections, csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

KEEP = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'sm__cycles_elapsed.avg.per_second', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum', 'sass__inst_executed_shared_loads',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio']

def raw(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]

# what was captured (so that DRAM traffic is compared with the algorithmic bytes of THAT launch)
CAPTURES = {
    "r2": {
        "backprojection": {"what": "second launch of scripts/quick_bench.py --det 2048 --vol 1024 --proj 480 --batch 256: 240 "
                                   "projections of 2048^2 into the 1024^3 volume (config 3 geometry)",
                           "algorithmic_bytes": 240 * 2048 * 2048 * 4 + 2 * 1024 ** 3 * 4},
        "fused": {"what": "one launch of the fused weight+filter kernel over 256 projections of 2048^2 (N = 4096), one "
                          "transform per 256-thread CTA (filter_wide = 0, the shape before the last kernel change)",
                  "algorithmic_bytes": 8 * 2048 * 2048 * 256},
        "fused-wide": {"what": "the same launch with two transforms per 512-thread CTA (filter_wide = 1, the default), "
                               "scripts/quick_bench.py --det 2048 --vol 256 --proj 256 --batch 256 --reps 1",
                       "algorithmic_bytes": 8 * 2048 * 2048 * 256},
    },
}
summary = {}
lines = [f"# ncu summaries, round {tag}", "",
         "Captured with `ncu --set full --clock-control none --import-source on` on a B200 through gpurun, after the",
         "same command had exited 0 without ncu.  Durations under ncu are serialised and cold-cache: use the bench's",
         "CUDA-event numbers for speed, these for WHERE the time goes.", ""]
for name, rep in (("backprojection (bp_tma_kernel)", f"gpurun_out/bp_{tag}_final.ncu-rep"),
                  ("fused weight+filter (filter_kernel)", f"gpurun_out/filter_{tag}_final.ncu-rep"),
                  ("fused-wide weight+filter, two transforms per CTA (filter_kernel<12, 1, 1>)", f"gpurun_out/filter_{tag}_wide.ncu-rep")):
    path = os.path.join(ROOT, rep)
    if not os.path.exists(path):
        continue
    hdr, units, data = raw(path)
    r = data[0]
    kn = r[hdr.index('Kernel Name')]
    lines += [f"## {name}", "", f"`{kn[:150]}`", "", "| metric | value | unit |", "|---|---|---|"]
    rec = {}
    for m in KEEP:
        if m in hdr:
            i = hdr.index(m)
            lines.append(f"| {m} | {r[i]} | {units[i]} |")
            try:
                rec[m] = float(r[i].replace(',', ''))
            except ValueError:
                rec[m] = r[i]
            rec[m + "__unit"] = units[i]
    lines.append("")
    cap = CAPTURES.get(tag, {}).get(name.split()[0])
    if cap:
        rec["capture"] = cap
        lines += [f"Captured launch: {cap['what']}; algorithmic HBM bytes {cap['algorithmic_bytes']:.3e}.", ""]
    summary[name.split()[0]] = rec

launch_csv = os.path.join(ROOT, f"gpurun_out/launches_{tag}.csv")
if os.path.exists(launch_csv):
    rows = [r for r in csv.reader(open(launch_csv)) if len(r) > 5]
    h = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    cols, data = rows[h], rows[h + 1:]
    ki, vi = cols.index('Kernel Name'), cols.index('Metric Value')
    agg = collections.OrderedDict()
    for r in data:
        a = agg.setdefault(r[ki].split('(')[0][:80], [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(',', ''))
    tot = sum(v[1] for v in agg.values())
    lines += ["## launch list of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline` (first launches)", "",
              "`ncu --metrics gpu__time_duration.sum --clock-control none -c 400/600` -- shares, not absolutes.", "",
              "| kernel | launches | total ms | share | avg us |", "|---|---|---|---|---|"]
    for k, v in agg.items():
        lines.append(f"| {k} | {v[0]} | {v[1]/1e6:.3f} | {v[1]/tot*100:.1f}% | {v[1]/v[0]/1e3:.1f} |")
    lines.append("")
    summary["launch_shares"] = {k: {"launches": v[0], "total_ms": v[1] / 1e6, "share": v[1] / tot} for k, v in agg.items()}
    import shutil
    shutil.copy(launch_csv, os.path.join(out_dir, f"{tag}_launches.csv"))

open(os.path.join(out_dir, f"{tag}_ncu_summary.md"), "w").write("\n".join(lines))
json.dump(summary, open(os.path.join(out_dir, f"{tag}_ncu_summary.json"), "w"), indent=1)
print("\n".join(lines))
