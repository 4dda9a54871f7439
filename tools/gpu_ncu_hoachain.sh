#!/bin/bash
# one full capture of each kernel of the configs[2] chain step (spatial decode, NFC / render, limiter, packer)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'limiter_kernel|hoasp_render|hoa_render_nfc|pcm_pack4' -s 20 -c 4 \
  -o gpurun_out/prof_hoachain -f \
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 3 --chain-steps 12 --chain-e2e-steps 3 --only hoa > gpurun_out/ncu_hoachain.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_hoachain.log | cut -c1-200
