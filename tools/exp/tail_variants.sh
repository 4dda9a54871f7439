#!/bin/bash
# A/B of the limiter variants (MPEGHB200_LIM_VARIANT: bit 0 = 256 threads per stream, bit 1 = release fast path of the
# smoother), on the chains of configs[2] and configs[4].
mkdir -p gpurun_out
run() {
  env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 3 --only objects,hoa \
      --chain-steps 300 --chain-e2e-steps 3 2> gpurun_out/variant.err |
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$*', {k: round(v['ms_per_step'] * 1000, 1) for k, v in d['configs'].items()})"
}
run MPEGHB200_LIM_VARIANT=0
run MPEGHB200_LIM_VARIANT=1
run MPEGHB200_LIM_VARIANT=2
run MPEGHB200_LIM_VARIANT=3

run MPEGHB200_LIM_VARIANT=3 MPEGHB200_CHAIN_FUSED_TAIL=1
timeout 900 python -m pytest tests/test_tail_gpu.py tests/test_chain_gpu.py tests/test_obj_render_gpu.py -x -q -m gpu 2>&1 | tail -3
MPEGHB200_LIM_VARIANT=0 timeout 900 python -m pytest tests/test_tail_gpu.py tests/test_chain_gpu.py tests/test_obj_render_gpu.py -x -q -m gpu 2>&1 | tail -3
