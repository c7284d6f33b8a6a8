#!/bin/bash
# round 2 (4 GPUs): C4 strong scaling point at N = 4 with the final code (coarse backward schedule, parity check included)
mkdir -p gpurun_out
export AMPNET_B200_DIST_TIMING=1
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/m_c4_4gpu.json 2> gpurun_out/m_c4_4gpu.err; echo "exit=$?" >> gpurun_out/m_c4_4gpu.err
grep -E "phase ms|exit=|Error|error" gpurun_out/m_c4_4gpu.err | tail -n 5
