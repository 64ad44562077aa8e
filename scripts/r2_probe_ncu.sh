#!/bin/bash
# round-2 ncu evidence (one GPU; every ncu command after the same command exited 0 without ncu):
#   launch list of `bench.py --steps 2 --warmup 3 --no-cpu-baseline` (config 3), --set full of the two hot kernels
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
Q="python scripts/quick_bench.py --det 2048 --vol 1024 --proj 480 --batch 256 --reps 2"
$B > gpurun_out/r2_ncu_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2.csv $B > gpurun_out/r2_ncu_launches.log 2>&1
$Q > gpurun_out/r2_ncu_plain_quick.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:bp_tma_kernel -s 1 -c 1 -o gpurun_out/bp_r2_final -f $Q > gpurun_out/r2_ncu_bp.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:filter_kernel -s 2 -c 1 -o gpurun_out/filter_r2_final -f $Q > gpurun_out/r2_ncu_filter.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_r2.csv
