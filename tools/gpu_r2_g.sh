#!/bin/bash
# round 2 (8 GPUs): C4 strong scaling with the ring-phased peer exchange, the NCCL all-to-all transport beside it, C5 at half scale
mkdir -p gpurun_out
export AMPNET_B200_DIST_TIMING=1
run() { # name, extra args...
  name=$1; shift
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 "$@" > gpurun_out/g_${name}.json 2> gpurun_out/g_${name}.err; echo "${name} exit=$?" >> gpurun_out/g_${name}.err
}
run c4_peer --steps 5 --warmup 3
run c5_peer --workload C5 --steps 3 --warmup 3 --no-parity-check
for n in c4_peer c5_peer; do grep -E "phase ms|exit=|Error|error|OutOfMemory" gpurun_out/g_${n}.err | tail -n 4; done
