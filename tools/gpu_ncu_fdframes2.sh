mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:fd_synth_ws_frames -s 6 -c 1 -o gpurun_out/prof_fdframes2 -f \
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs --e2e-steps 3 > gpurun_out/ncu_fdframes2.log 2>&1
echo "ncu rc=$?"
