#!/bin/bash
mkdir -p gpurun_out
for side in long mixed; do
for v in ${VARIANTS:-0 3 5}; do
  MPEGHB200_BENCH_SIDE=$side MPEGHB200_FD_VARIANT=$v python bench.py --steps 200 --warmup 20 --no-cpu-baseline --e2e-steps 3 > gpurun_out/bench_${side}_v$v.log 2>&1
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${side}_v$v.log").read().strip().splitlines()[-1])
    print("$side variant $v ms/step", d["ms_per_step"], "frac", d["roofline"]["frac"])
except Exception as e:
    print("variant $v failed", e)
PY
done; done
