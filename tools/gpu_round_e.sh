#!/bin/bash
mkdir -p gpurun_out
for v in ${VARIANTS:-7 8}; do
  MPEGHB200_FD_VARIANT=$v timeout 180 python -m pytest tests/test_fd_synth_gpu.py -x -q -m gpu > gpurun_out/pytest_v$v.log 2>&1; echo "variant $v pytest rc=$?"; tail -2 gpurun_out/pytest_v$v.log
  for mode in base; do
  MPEGHB200_FD_VARIANT=$v timeout 120 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --e2e-steps 3 > gpurun_out/bench_v$v.log 2>&1
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_v$v.log").read().strip().splitlines()[-1])
    print("variant $v ms/step", d["ms_per_step"], "frac", d["roofline"]["frac"], "mixed", d["mixed_sequence_variant"]["ms_per_step"], d["mixed_sequence_variant"]["roofline_frac"], "e2e ms", d["e2e"]["ms_per_step"])
except Exception as e:
    print("variant $v failed", e)
PY
  done
done
