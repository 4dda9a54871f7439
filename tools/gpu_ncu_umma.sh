#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/bench_stages.py hoa_render > gpurun_out/umma_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:hoa_render_umma -s 3 -c 1 -o gpurun_out/prof_umma -f \
  python tools/bench_stages.py hoa_render > gpurun_out/ncu_umma.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_umma.log | cut -c1-200
