#!/bin/bash
# One GPU-box session: GPU tests, phase timers, bench, ncu launch list and a full capture of the attention kernels.
# usage: tools/gpu_round.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$tag.log
python tools/phase_profile.py > gpurun_out/phase_$tag.log 2>&1; echo "phase rc=$?"
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; cat gpurun_out/bench_$tag.json
python bench.py --workload C4s --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --workload C4s --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu1_$tag.log 2>&1; echo "ncu1 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:attn_ -s 9 -c 3 -o gpurun_out/prof_attn_$tag \
    python bench.py --workload C4s --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu2_$tag.log 2>&1; echo "ncu2 rc=$?"
ls -la gpurun_out | tail -8
