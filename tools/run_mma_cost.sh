#!/bin/bash
# tcgen05.mma issue / completion cost sweep (umma_probe mode 7): TS|SS x N x #accumulators
mkdir -p gpurun_out
{
for ss in 0 1; do
  for cfg in "16 1" "16 2" "16 4" "16 8" "32 1" "32 4" "64 1" "64 4" "128 1" "128 2" "256 1"; do
    timeout 60 ./tools/umma_probe 7 $ss $cfg | grep RESULT
  done
done
} > gpurun_out/mma_cost.log 2>&1
cat gpurun_out/mma_cost.log
