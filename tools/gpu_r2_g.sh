#!/bin/bash
# round 2 (8 GPUs): C4 strong scaling with the ring-phased peer exchange (parity check against the oracle included)
mkdir -p gpurun_out
export AMPNET_B200_DIST_TIMING=1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/g_c4_peer.json 2> gpurun_out/g_c4_peer.err; echo "c4_peer exit=$?" >> gpurun_out/g_c4_peer.err
grep -E "phase ms|exit=|Error|error|OutOfMemory" gpurun_out/g_c4_peer.err | tail -n 6
