#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:limiter_gains -s 8 -c 1 -o gpurun_out/prof_lim -f \
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 3 --chain-steps 6 --chain-e2e-steps 3 --only hoa > gpurun_out/ncu_lim.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_lim.log | cut -c1-200
