#!/bin/bash
# One GPU-box session: GPU tests, smoke(), bench, ncu launch list and full captures of the attention kernels
# (C4s = 1/10-scale workload for the launch list / source-level capture, C4 for the per-launch DRAM traffic).
# usage: tools/gpu_round.sh <tag>        (tools/phase_profile.py is run separately: PHASES=1 tools/gpu_round.sh <tag>)
tag=${1:-x}
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$tag.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_$tag.log
if [ -n "$PHASES" ]; then timeout 120 python tools/phase_profile.py > gpurun_out/phase_$tag.log 2>&1; echo "phase rc=$?"; fi
timeout 240 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; cat gpurun_out/bench_$tag.json
timeout 120 python bench.py --workload C4s --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_$tag.log 2>&1 &&
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --workload C4s --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu1_$tag.log 2>&1; echo "ncu1 rc=$?"
timeout 240 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 9 -c 3 -o gpurun_out/prof_attn_$tag \
    python bench.py --workload C4s --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu2_$tag.log 2>&1; echo "ncu2 rc=$?"
timeout 240 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:attn_ -s 9 -c 3 --csv \
    --log-file gpurun_out/traffic_c4_$tag.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu3_$tag.log 2>&1; echo "ncu3 rc=$?"
ls -la gpurun_out | tail -8
