#!/bin/bash
# A/B of the limiter launch forms (MPEGHB200_LIM_VARIANT: bit 0 = 256 threads per stream, bit 1 = release fast path of
# the smoother, bit 2 = batched channel loads for launches of at most four streams per SM) on the chains of configs[2]
# and configs[4], then the parity suites of the tail.  Measured on B200 (us per chain step, objects / hoa):
#   0: 161.7 / 393.0   1: 159.4 / 398.9   2: 155.6 / 399.3   3: 152.3 / 399.6   7 (default): 148.1 / 401.9
mkdir -p gpurun_out
run() {
  env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 3 --only objects,hoa \
      --chain-steps 300 --chain-e2e-steps 3 2> gpurun_out/variant.err |
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$*', {k: round(v['ms_per_step'] * 1000, 1) for k, v in d['configs'].items()})"
}
for v in 0 1 2 3 7; do run MPEGHB200_LIM_VARIANT=$v; done
timeout 900 python -m pytest tests/test_tail_gpu.py tests/test_chain_gpu.py tests/test_obj_render_gpu.py -x -q -m gpu 2>&1 | tail -3
