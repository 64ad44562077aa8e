#!/bin/bash
# round-2 probe 6 (one GPU): the whole GPU suite on the float-row kernel, bench lines for config 3 and 2
mkdir -p gpurun_out
rm -f gpurun_out/r2_parity.jsonl
( time timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "parity|passed|failed|Error|error|GUPS" | tail -150 ) > gpurun_out/r2_p6_suite.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_p6_bench_c3.json 2> gpurun_out/r2_p6_bench_c3.err
timeout 300 python bench.py --config c2 --steps 5 --warmup 3 > gpurun_out/r2_p6_bench_c2.json 2> gpurun_out/r2_p6_bench_c2.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2_p6_ref_c3.json 2> gpurun_out/r2_p6_ref_c3.err
