#!/bin/bash
# round-2 probe 5 (one GPU): backprojection with the reference's float row arithmetic -- parity and speed
mkdir -p gpurun_out
( timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cases.py tests/test_gpu_group.py -m gpu -q -x 2>&1 | tail -6 ) > gpurun_out/r2_p5_tests.log 2>&1
echo "== c2" > gpurun_out/r2_p5_bench.log
timeout 200 python scripts/quick_bench.py --batch 256 --reps 4 --check >> gpurun_out/r2_p5_bench.log 2>&1
echo "== c3" >> gpurun_out/r2_p5_bench.log
timeout 300 python scripts/quick_bench.py --det 2048 --vol 1024 --proj 1440 --batch 256 --reps 3 >> gpurun_out/r2_p5_bench.log 2>&1
echo "== natural 1024" >> gpurun_out/r2_p5_bench.log
timeout 300 python scripts/quick_bench.py --det 1024 --proj 360 --natural --batch 256 --reps 3 --check >> gpurun_out/r2_p5_bench.log 2>&1
