#!/bin/bash
# N=8 with mirror-balanced boxes: two z-runs (each holds one of the slow outermost 128-slice blocks) x four x-parts
mkdir -p gpurun_out
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517"
PARIS_B200_GROUP_TRACE=1 timeout 300 $R bench.py --gpus 8 --steps 5 --warmup 3 --x-parts 4 > gpurun_out/r2_n8b_c3.json 2> gpurun_out/r2_n8b_c3.err; echo "c3 rc=$?" > gpurun_out/r2_n8b_rc.log
timeout 200 $R bench.py --gpus 8 --steps 5 --warmup 3 --config c2 --x-parts 4 > gpurun_out/r2_n8b_c2.json 2> gpurun_out/r2_n8b_c2.err; echo "c2 rc=$?" >> gpurun_out/r2_n8b_rc.log
