"""gpurun_out/r2_parity.jsonl (written by tests/test_gpu_fullsize.py on a B200) -> profiles/r2_parity.json:
every measured (max, rmse) / C per configuration and check, the worst per configuration, the margins."""
import collections, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = os.path.join(ROOT, "gpurun_out", "r2_parity.jsonl")
recs = collections.OrderedDict()
for line in open(src):
    r = json.loads(line)
    recs[(r["config"], r["check"])] = r            # a later run of the same check replaces the earlier one
by_cfg = collections.OrderedDict()
for (cfg, _), r in recs.items():
    by_cfg.setdefault(cfg, []).append(r)
out = {"tolerance": {"max_over_C": 1e-4, "rmse_over_C": 1e-5, "C": "delta_mu * N_proj / (8 pi) (SURVEY F6)"},
       "source": "tests/test_gpu_fullsize.py on a B200 (gpurun); oracle = oracle/fdk_oracle.c, bit-identical to the "
                 "reference's OpenMP backend (tests/test_oracle.py, tests/test_oracle_blocks.py)",
       "configs": collections.OrderedDict()}
for cfg, rs in by_cfg.items():
    oracle = [r for r in rs if r["check"].startswith("oracle")]
    exact = [r for r in rs if not r["check"].startswith("oracle")]
    entry = {"checks": rs}
    if oracle:
        entry["worst_vs_oracle"] = {"max_over_C": max(r["max_over_C"] for r in oracle), "rmse_over_C": max(r["rmse_over_C"] for r in oracle)}
    if exact:
        entry["worst_vs_exact_kernel"] = {"max_over_C": max(r["max_over_C"] for r in exact), "rmse_over_C": max(r["rmse_over_C"] for r in exact)}
    worst = max(r["max_over_C"] for r in rs)
    entry["share_of_max_budget_used"] = worst / 1e-4
    out["configs"][cfg] = entry
json.dump(out, open(os.path.join(ROOT, "profiles", "r2_parity.json"), "w"), indent=1)
for cfg, e in out["configs"].items():
    print(cfg, {k: v for k, v in e.items() if k != "checks"})
