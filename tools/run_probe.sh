#!/bin/bash
# Runs the tcgen05 probe variants, one process each (a faulting variant cannot poison the others).
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
for v in "0 0" "0 1" "0 3" "4 2" "3" \
         "1 0 16 1024 1" "1 1 16 1024 1" "1 3 16 1024 1" "1 1 1024 16 1" "1 1 128 1024 1" "1 1 16 1024 0" \
         "2 0 16 1024 1" "2 2 16 1024 1" "6 0 16384 1024" "6 0 1024 16384" "5"; do
  echo "--- umma_probe $v"
  timeout 60 ./tools/umma_probe $v
  echo "exit=$?"
done
} > gpurun_out/probe.log 2>&1
tail -60 gpurun_out/probe.log
