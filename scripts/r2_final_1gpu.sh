#!/bin/bash
# final one-GPU pass of round 2: the GPU suite, the bench lines, the per-slab backprojection times
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_final_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/r2_final_rc.log
timeout 400 python bench.py > gpurun_out/r2_final_c3.json 2> gpurun_out/r2_final_c3.err; echo "c3 rc=$?" >> gpurun_out/r2_final_rc.log
timeout 300 python bench.py --config c2 > gpurun_out/r2_final_c2.json 2> gpurun_out/r2_final_c2.err; echo "c2 rc=$?" >> gpurun_out/r2_final_rc.log
timeout 300 python scripts/r2_slab_times.py > gpurun_out/r2_slab_times.json 2> gpurun_out/r2_slab_times.err; echo "slabs rc=$?" >> gpurun_out/r2_final_rc.log
timeout 500 python bench.py --samples u16 --no-cpu-baseline > gpurun_out/r2_final_c3_u16.json 2> gpurun_out/r2_final_c3_u16.err; echo "u16 rc=$?" >> gpurun_out/r2_final_rc.log
timeout 300 python bench.py --impl reference > gpurun_out/r2_final_reference.json 2> gpurun_out/r2_final_reference.err; echo "ref rc=$?" >> gpurun_out/r2_final_rc.log
