#!/bin/bash
# round-2 probe 2: whole GPU suite with the full-size oracle-block tests, then the default bench (config 3)
mkdir -p gpurun_out
rm -f gpurun_out/r2_parity.jsonl
( time python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -120 ) > gpurun_out/r2_p2_pytest.log 2>&1
python bench.py --steps 3 --warmup 3 > gpurun_out/r2_p2_bench_c3.json 2> gpurun_out/r2_p2_bench_c3.err
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r2_p2_ref_c3.json 2> gpurun_out/r2_p2_ref_c3.err
