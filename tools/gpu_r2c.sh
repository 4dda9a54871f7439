#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tail_gpu.py tests/test_chain_gpu.py tests/test_canary_gpu.py -x -q -m gpu > gpurun_out/r2c_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2c_tests.log
for f in 1 0; do
MPEGHB200_CHAIN_FUSED_TAIL=$f timeout 900 python bench.py --steps 200 --warmup 20 --e2e-steps 20 --no-cpu-baseline --only hoa,objects > gpurun_out/r2c_bench_$f.json 2> gpurun_out/r2c_bench.err; echo "bench fused=$f rc=$?"; tail -3 gpurun_out/r2c_bench.err
python - <<PY
import json
l=json.load(open('gpurun_out/r2c_bench_$f.json'))
for k,v in l['configs'].items(): print(k,'ms',round(v['ms_per_step'],4),'frac',round(v['roofline']['frac'],3),'launches/step',v['launches_per_step'],'e2e ms',round(v['e2e']['ms_per_step'],3))
PY
done
