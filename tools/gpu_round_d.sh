#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -2 gpurun_out/pytest_gpu.log
for side in long mixed; do
for v in ${VARIANTS:-0 2 3 6}; do
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
