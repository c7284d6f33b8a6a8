#!/bin/bash
# round 2: ncu launch list and one --set full capture of the attention kernels (C4s = C4 token shape at 1/10 scale)
mkdir -p gpurun_out
CMD="python bench.py --workload C4s --steps 2 --warmup 3 --no-cpu-baseline --no-parity-check"
timeout 200 $CMD > gpurun_out/k_plain.json 2> gpurun_out/k_plain.err &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/k_launches.csv $CMD > gpurun_out/k_ncu1.log 2>&1
echo "launch list exit=$?" >> gpurun_out/k_ncu1.log
timeout 200 $CMD > gpurun_out/k_plain2.json 2> gpurun_out/k_plain2.err &&
timeout 500 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 9 -c 3 -o gpurun_out/k_prof -f $CMD > gpurun_out/k_ncu2.log 2>&1
echo "full capture exit=$?" >> gpurun_out/k_ncu2.log
tail -n 3 gpurun_out/k_ncu1.log gpurun_out/k_ncu2.log; ls -la gpurun_out/k_*
