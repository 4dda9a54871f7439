mkdir -p gpurun_out
[ -x tools/exp/fp_probe ] && tools/exp/fp_probe > gpurun_out/fp_probe.txt 2>&1; cat gpurun_out/fp_probe.txt | head -20
ncu --set full --clock-control none --import-source on -k regex:hoa_render -s 6 -c 2 -o gpurun_out/prof_hoaren -f \
  python tools/bench_stages.py hoa_render > gpurun_out/ncu_hoaren.log 2>&1
echo rc=$?
