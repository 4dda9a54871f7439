python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_r1m.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu_r1m.log
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py --steps 2000 --warmup 50 > gpurun_out/bench_r1m.json 2> gpurun_out/bench_r1m.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench_r1m.json
