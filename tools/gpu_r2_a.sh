#!/bin/bash
# round 2, call A: GPU tests on the untouched kernels + new host logic, MUFU probe, C4 and C5s (head-group decomposition) baselines
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.log
timeout 60 ./tools/mufu_probe > gpurun_out/a_mufu.log 2>&1; echo "mufu exit=$?" >> gpurun_out/a_mufu.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest exit=$?" >> gpurun_out/a_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/a_bench_c4.json 2> gpurun_out/a_bench_c4.err; echo "bench c4 exit=$?" >> gpurun_out/a_bench_c4.err
timeout 600 python bench.py --workload C5s --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench_c5s.json 2> gpurun_out/a_bench_c5s.err; echo "bench c5s exit=$?" >> gpurun_out/a_bench_c5s.err
tail -5 gpurun_out/a_mufu.log gpurun_out/a_pytest.log gpurun_out/a_bench_c4.err gpurun_out/a_bench_c5s.err
