#!/bin/bash
# one full capture of the binaural block kernel inside the configs[3] chain step
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'binaural_block' -s 2 -c 1 -o gpurun_out/prof_binchain -f \
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 3 --chain-steps 6 --chain-e2e-steps 3 --only binaural > gpurun_out/ncu_binchain.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_binchain.log | cut -c1-200
