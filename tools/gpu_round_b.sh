#!/bin/bash
mkdir -p gpurun_out
./tools/exp/copy_probe > gpurun_out/copy_probe.log 2>&1; cat gpurun_out/copy_probe.log
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke 2>&1 | tail -2
for v in ${VARIANTS:-0 1 2 3 4 5 6}; do
  MPEGHB200_FD_VARIANT=$v python bench.py --steps 200 --warmup 20 --no-cpu-baseline --e2e-steps 3 > gpurun_out/bench_v$v.log 2>&1
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_v$v.log").read().strip().splitlines()[-1])
    print("variant $v ms/step", d["ms_per_step"], "frac", d["roofline"]["frac"])
except Exception as e:
    print("variant $v failed", e)
PY
done
