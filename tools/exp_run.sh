#!/bin/bash
# runs bench.py against every experimental library under exp/ and prints the kernel timings
mkdir -p gpurun_out
: > gpurun_out/exp_results.txt
for d in exp/*/; do
  n=$(basename $d)
  FL_LIB=$PWD/$d/lib.so python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/exp_$n.json 2> gpurun_out/exp_$n.err
  python - "$n" gpurun_out/exp_$n.json >> gpurun_out/exp_results.txt <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print("%-24s step %.4f ms  derivs %.4f ms  values+assemble %.4f ms  e2e %.4f ms" % (sys.argv[1], j["ms_per_step"], j["kernels_ms"]["node_derivs"], j["kernels_ms"]["values+assemble"], j["e2e"]["ms_per_step"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
cat gpurun_out/exp_results.txt
