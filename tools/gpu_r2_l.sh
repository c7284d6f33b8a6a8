#!/bin/bash
# round 2: DRAM traffic of the three attention kernels at the full C4 size (one ncu --set full capture, three launches)
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity-check"
timeout 200 $CMD > gpurun_out/l_plain.json 2> gpurun_out/l_plain.err &&
timeout 500 ncu --set full --clock-control none -k regex:attn_ -s 9 -c 3 -o gpurun_out/l_prof -f $CMD > gpurun_out/l_ncu.log 2>&1
echo "full capture exit=$?" >> gpurun_out/l_ncu.log
tail -n 3 gpurun_out/l_ncu.log
