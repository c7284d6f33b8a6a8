#!/bin/bash
# round 2, call E (1 GPU): head_dim 8 through cp.async-padded tiles, ring-phase kernels (single d_qkv buffer), benches
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_zhd8.py tests/test_gpu_ring_phases.py tests/test_gpu_halo.py -q > gpurun_out/e_hd8.log 2>&1; echo "hd8 exit=$?" >> gpurun_out/e_hd8.log
timeout 600 python bench.py --workload C5s --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/e_bench_c5s.json 2> gpurun_out/e_bench_c5s.err; echo "bench c5s exit=$?" >> gpurun_out/e_bench_c5s.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/e_bench_c4.json 2> gpurun_out/e_bench_c4.err; echo "bench c4 exit=$?" >> gpurun_out/e_bench_c4.err
grep -E "passed|failed|exit=" gpurun_out/e_hd8.log | tail -n 5; tail -n 3 gpurun_out/e_bench_c5s.err; tail -n 2 gpurun_out/e_bench_c4.err
