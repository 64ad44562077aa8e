#!/bin/bash
# run the bench at N = $@ (default 1 2 4 8) back to back, one JSON line each -> gpurun_out/scale_N.json
mkdir -p gpurun_out
for n in "${@:-1 2 4 8}"; do
  if [ "$n" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  fi
  python - <<PY
import json
try:
    d = [json.loads(l) for l in open("gpurun_out/scale_$n.json") if l.startswith("{")][-1]
    print("N=$n", "ms/step", round(d["ms_per_step"], 2), "GUPS", round(d["value"], 1), "stages", {k: round(v, 2) for k, v in d["stage_ms"].items() if k != "note"}, "e2e ms", round(d["e2e"]["seconds"] * 1e3, 1))
except Exception as e:
    print("N=$n failed", e); print(open("gpurun_out/scale_$n.err").read()[-1500:])
PY
done
