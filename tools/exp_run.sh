#!/bin/bash
# runs bench.py against every experimental library under exp/ and prints the kernel timings (8000-node lap and a 1024-lap batch)
mkdir -p gpurun_out
: > gpurun_out/exp_results.txt
for d in exp/*/; do
  n=$(basename $d)
  [ -f $d/lib.so ] || continue
  FL_LIB=$PWD/$d/lib.so python bench.py --steps 50 --warmup 10 --no-cpu-baseline --laps ${EXP_LAPS:-1024} --solve-laps 0 --gg-speeds 0 --single-laps 0 > gpurun_out/exp_$n.json 2> gpurun_out/exp_$n.err
  python - "$n" gpurun_out/exp_$n.json >> gpurun_out/exp_results.txt <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    b = j.get("batched", {})
    print("%-24s lap8000: step %.4f ms derivs %.4f values+assemble %.4f | batch: step %.3f ms values %.3f derivs %.3f frac %.3f" % (
        sys.argv[1], j["ms_per_step"], j["kernels_ms"]["node_derivs"], j["kernels_ms"]["values+assemble"],
        b.get("ms_per_step", -1), b.get("kernels_ms", {}).get("node_values", -1), b.get("kernels_ms", {}).get("node_derivs+assemble", -1), b.get("roofline", {}).get("frac", -1)))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
cat gpurun_out/exp_results.txt
