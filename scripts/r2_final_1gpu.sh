#!/bin/bash
# final one-GPU pass of round 2: the GPU suite, smoke, the bench line, A/B of the filter's CTA shape, 16-bit samples
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_final_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/r2_final_rc.log
timeout 60 python -c 'import __graft_entry__ as g; g.smoke(); print("smoke ok")' > gpurun_out/r2_final_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_final_rc.log
timeout 200 python bench.py > gpurun_out/r2_final_c3.json 2> gpurun_out/r2_final_c3.err; echo "c3 rc=$?" >> gpurun_out/r2_final_rc.log
PARIS_B200_FILTER_WIDE=0 timeout 150 python bench.py --steps 3 --no-cpu-baseline > gpurun_out/r2_final_c3_narrow.json 2> gpurun_out/r2_final_c3_narrow.err; echo "narrow rc=$?" >> gpurun_out/r2_final_rc.log
timeout 200 python bench.py --samples u16 --steps 3 --no-cpu-baseline > gpurun_out/r2_final_c3_u16.json 2> gpurun_out/r2_final_c3_u16.err; echo "u16 rc=$?" >> gpurun_out/r2_final_rc.log
