#!/bin/bash
# usage: gpu_ncu.sh <tag> [variant]   -- plain run first, then one full ncu capture of the FD kernel
TAG=$1; V=${2:-0}
mkdir -p gpurun_out
export MPEGHB200_FD_VARIANT=$V
export MPEGHB200_BENCH_SIDE=${SIDE:-mixed}
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fd_synth -s 4 -c 1 -o gpurun_out/prof_$TAG -f \
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/ncu_$TAG.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_$TAG.log | cut -c1-300
