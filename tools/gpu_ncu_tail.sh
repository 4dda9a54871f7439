#!/bin/bash
# one ncu --set full capture each of the limiter and the PCM packer inside the HOA chain (1024 streams x 24 channels)
mkdir -p gpurun_out
for k in limiter_kernel pcm_pack_kernel; do
ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -o gpurun_out/prof_$k -f \
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 3 --chain-steps 6 --chain-e2e-steps 3 --only hoa > gpurun_out/ncu_$k.log 2>&1
echo "$k rc=$?"
done
