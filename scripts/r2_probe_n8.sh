#!/bin/bash
# round-2, eight GPUs: scaling lines for config 3 (north star), 2, 5 (two slabs per GPU), 4; config 3 at N=4
mkdir -p gpurun_out
export PARIS_B200_GROUP_TIMEOUT_S=30
nvidia-smi topo -m > gpurun_out/r2_n8_topo.log 2>&1
free -g >> gpurun_out/r2_n8_topo.log; df -h /dev/shm >> gpurun_out/r2_n8_topo.log; nproc >> gpurun_out/r2_n8_topo.log
run() { # name, ranks, args...
  name=$1; n=$2; shift; shift
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $n "$@" > gpurun_out/r2_n8_$name.json 2> gpurun_out/r2_n8_$name.err
  echo "$name rc=$?" >> gpurun_out/r2_n8_rc.log
}
PARIS_B200_GROUP_TRACE=1 run c3 8 --steps 4 --warmup 3
PARIS_B200_GROUP_TRACE=1 run c2 8 --config c2 --steps 6 --warmup 3
run c3_n4 4 --steps 3 --warmup 3
run c5 8 --config c5 --steps 2 --warmup 3
run c4 8 --config c4 --steps 2 --warmup 3
