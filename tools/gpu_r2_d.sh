#!/bin/bash
# round 2 (2 GPUs): partitioned path against the oracle (both transports), then C4 strong scaling at N = 2
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/d_topo.log 2>&1
timeout 420 python -m pytest tests/test_gpu_distributed.py -x -q > gpurun_out/d_dist.log 2>&1; echo "dist exit=$?" >> gpurun_out/d_dist.log
export AMPNET_B200_DIST_TIMING=1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/d_bench_2gpu.json 2> gpurun_out/d_bench_2gpu.err; echo "bench peer exit=$?" >> gpurun_out/d_bench_2gpu.err
grep -E "passed|failed|exit=|Error" gpurun_out/d_dist.log | tail -n 12; grep -E "phase ms|exit=|Error|error" gpurun_out/d_bench_2gpu.err | tail -n 12
