#!/bin/bash
# round 2, first GPU pass: chain tests first (new code), the whole GPU suite, smoke, bench (both arms)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_chain_gpu.py -x -q -m gpu > gpurun_out/r2a_chain.log 2>&1; echo "chain pytest rc=$?"; tail -15 gpurun_out/r2a_chain.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2a_pytest.log
python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py --steps 2000 --warmup 50 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/r2a_bench.err; cut -c1-300 gpurun_out/r2a_bench.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/r2a_bench_ref.json
nproc; lscpu | grep "Model name"
