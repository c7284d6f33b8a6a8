#!/bin/bash
# round 2 (1 GPU): whole GPU suite on the final kernels (bf16 gradient rows, coarse backward, head_dim 8), smoke, C4 + C5s bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/j_pytest.log 2>&1; echo "pytest exit=$?" >> gpurun_out/j_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/j_smoke.log 2>&1; echo "smoke exit=$?" >> gpurun_out/j_smoke.log
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/j_bench_c4.json 2> gpurun_out/j_bench_c4.err; echo "bench c4 exit=$?" >> gpurun_out/j_bench_c4.err
timeout 400 python bench.py --workload C5s --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/j_bench_c5s.json 2> gpurun_out/j_bench_c5s.err; echo "bench c5s exit=$?" >> gpurun_out/j_bench_c5s.err
tail -n 6 gpurun_out/j_pytest.log; tail -n 3 gpurun_out/j_smoke.log; tail -n 2 gpurun_out/j_bench_c4.err gpurun_out/j_bench_c5s.err
