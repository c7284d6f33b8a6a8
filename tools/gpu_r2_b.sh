#!/bin/bash
# round 2, call B (1 GPU): ring-phase kernels of a virtual world + the whole GPU suite
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ring_phases.py -x -q > gpurun_out/b_ring.log 2>&1; echo "ring exit=$?" >> gpurun_out/b_ring.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/b_pytest.log 2>&1; echo "pytest exit=$?" >> gpurun_out/b_pytest.log
timeout 300 python bench.py --workload C4s --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_bench_c4s.json 2> gpurun_out/b_bench_c4s.err; echo "bench exit=$?" >> gpurun_out/b_bench_c4s.err
tail -n 30 gpurun_out/b_ring.log; tail -n 8 gpurun_out/b_pytest.log; tail -n 3 gpurun_out/b_bench_c4s.err
