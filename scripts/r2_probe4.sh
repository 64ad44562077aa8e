#!/bin/bash
# round-2 probe 4 (one GPU): filter with derived twiddles (tests + speed), backprojection with rotating producer (A/B),
# detector-size growth of the row-rounding error, config-2 bench line at N=1
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cases.py -m gpu -q -x 2>&1 | tail -5 ) > gpurun_out/r2_p4_tests.log 2>&1
for lib in "" "paris_b200/libparis_b200_rot.so"; do
  echo "== lib '$lib' c2" >> gpurun_out/r2_p4_ab.log
  PARIS_B200_LIB=$lib timeout 200 python scripts/quick_bench.py --batch 256 --reps 4 --check >> gpurun_out/r2_p4_ab.log 2>&1
  echo "== lib '$lib' c3" >> gpurun_out/r2_p4_ab.log
  PARIS_B200_LIB=$lib timeout 300 python scripts/quick_bench.py --det 2048 --vol 1024 --proj 1440 --batch 256 --reps 3 >> gpurun_out/r2_p4_ab.log 2>&1
  echo "== lib '$lib' natural 1024" >> gpurun_out/r2_p4_ab.log
  PARIS_B200_LIB=$lib timeout 300 python scripts/quick_bench.py --det 1024 --proj 360 --natural --batch 256 --reps 3 --check >> gpurun_out/r2_p4_ab.log 2>&1
done
( timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -s -k "growth or config3" 2>&1 | grep -E "parity|passed|failed|Error" | tail -40 ) > gpurun_out/r2_p4_growth.log 2>&1
timeout 300 python bench.py --config c2 --steps 5 --warmup 3 > gpurun_out/r2_p4_bench_c2.json 2> gpurun_out/r2_p4_bench_c2.err
