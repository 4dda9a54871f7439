#!/bin/bash
# bench headline (no chains), then one ncu --set full capture of the F-frame FD kernel and the launch list of the same command
mkdir -p gpurun_out
python bench.py --no-configs --no-cpu-baseline > gpurun_out/bench_fdframes.json 2> gpurun_out/bench_fdframes.err
echo "bench rc=$?"; cut -c1-1500 gpurun_out/bench_fdframes.json
ncu --set full --clock-control none --import-source on -k regex:fd_synth_ws_frames -s 6 -c 1 -o gpurun_out/prof_fdframes -f \
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs --e2e-steps 3 > gpurun_out/ncu_fdframes.log 2>&1
echo "ncu rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2e_fd_launches.csv \
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs --e2e-steps 3 > gpurun_out/ncu_fdl.log 2>&1
echo "launch list rc=$?"
