#!/bin/bash
# round-2 probe 3: group tests on one GPU, full-size oracle-block tests, bench N=1 through the new arm
mkdir -p gpurun_out
rm -f gpurun_out/r2_parity.jsonl
( time timeout 600 python -m pytest tests/test_gpu_group.py -m gpu -q -s -x 2>&1 | tail -40 ) > gpurun_out/r2_p3_group.log 2>&1
( time timeout 1500 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity.py -m gpu -q -s 2>&1 | grep -v "^$" | tail -150 ) > gpurun_out/r2_p3_fullsize.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_p3_bench_c3.json 2> gpurun_out/r2_p3_bench_c3.err
