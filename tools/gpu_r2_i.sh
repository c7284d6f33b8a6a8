#!/bin/bash
# A/B: share of the exponentials on the FMA pipe (compile-time variants of the attention kernels), C4 on one GPU
mkdir -p gpurun_out
for v in base poly_dq poly_dkv poly_both poly_fwdhalf; do
  lib=""; [ "$v" != base ] && lib="$PWD/ampnet_b200/libampconv_$v.so"
  AMPNET_B200_LIB=$lib timeout 240 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-parity-check > gpurun_out/i_$v.json 2> gpurun_out/i_$v.err; echo "$v exit=$?"
  python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/i_$v.json").read().strip().splitlines()[-1])
    print("$v", "ms/step %.2f" % j["ms_per_step"], {k: round(x, 2) for k, x in j["roofline"]["kernel_ms"].items()})
except Exception as e:
    print("$v", "no result", e)
PY
done > gpurun_out/i_summary.log 2>&1
cat gpurun_out/i_summary.log
