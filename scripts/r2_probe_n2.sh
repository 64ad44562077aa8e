#!/bin/bash
# round-2, two GPUs: group over CUDA IPC (tests), CLI on two devices, bench N=2 on config 3 and 2
mkdir -p gpurun_out
export PARIS_B200_GROUP_TIMEOUT_S=20
nvidia-smi -L > gpurun_out/r2_n2_gpus.log
nvidia-smi topo -m >> gpurun_out/r2_n2_gpus.log 2>&1
( time timeout 400 python -m pytest tests/test_gpu_group.py tests/test_gpu_cli.py -m gpu -q -s 2>&1 | tail -40 ) > gpurun_out/r2_n2_tests.log 2>&1
run() { # name, args...
  name=$1; shift
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus 2 "$@" > gpurun_out/r2_n2_$name.json 2> gpurun_out/r2_n2_$name.err
  echo "$name rc=$?" >> gpurun_out/r2_n2_rc.log
}
run c3 --steps 3 --warmup 3
run c2 --config c2 --steps 5 --warmup 3
run c3_kernel --steps 3 --warmup 3 --exchange kernel
run c2_whole --config c2 --steps 5 --warmup 3 --whole-projections
