#!/bin/bash
# round-2 probe 1: GPU suite after the ADVICE fixes, gather micro-benchmark, tile-numbering A/B
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_p1_pytest.log
./scripts/microbench/gather_peak > gpurun_out/r2_gather_peak.json 2> gpurun_out/r2_gather_peak.err
for sw in 0 16 8 32; do
  echo "== c2 swizzle $sw" >> gpurun_out/r2_p1_swizzle.log
  python scripts/quick_bench.py --batch 256 --reps 4 --swizzle $sw >> gpurun_out/r2_p1_swizzle.log 2>&1
done
for sw in 0 16; do
  echo "== c3 swizzle $sw" >> gpurun_out/r2_p1_swizzle.log
  python scripts/quick_bench.py --det 2048 --vol 1024 --proj 1440 --batch 256 --reps 3 --swizzle $sw >> gpurun_out/r2_p1_swizzle.log 2>&1
done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/r2_p1_smi.log
