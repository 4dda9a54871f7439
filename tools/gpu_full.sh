#!/bin/bash
# full round check: gpu tests, smoke, bench (both arms), ncu launch list + one full capture of the top kernel
TAG=${1:-r1}
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu_$TAG.log
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/bench_$TAG.json
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref_$TAG.json 2>&1; cut -c1-400 gpurun_out/bench_ref_$TAG.json
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/ncu_l_$TAG.log 2>&1
echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:fd_synth -s 4 -c 1 -o gpurun_out/prof_$TAG -f \
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/ncu_f_$TAG.log 2>&1
echo "ncu full rc=$?"
