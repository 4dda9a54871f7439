#!/bin/bash
# one full capture of each kernel of the configs[4] chain step (gains, mix, limiter, packer)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'limiter_|obj_mix|obj_gains|pcm_pack4' -s 20 -c 4 \
  -o gpurun_out/prof_objchain -f \
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 3 --chain-steps 12 --chain-e2e-steps 3 --only objects > gpurun_out/ncu_objchain.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_objchain.log | cut -c1-200
