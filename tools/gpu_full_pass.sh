#!/bin/bash
# full pass: whole GPU suite, smoke, bench both arms, launch list, one full capture of the headline kernel
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2i_pytest.log
python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2i_bench.err; cut -c1-200 gpurun_out/r2i_bench.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2i_bench_ref.json 2> gpurun_out/r2i_bench_ref.err; echo "ref rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 3 --chain-steps 6 --chain-e2e-steps 3 > gpurun_out/r2i_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2i_launches.csv \
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 3 --chain-steps 6 --chain-e2e-steps 3 > gpurun_out/r2i_ncu_l.log 2>&1
echo "ncu list rc=$?"
python tools/ncu_launch_avg.py gpurun_out/r2i_launches.csv | grep mpeghb200 | cut -c1-120
