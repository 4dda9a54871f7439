python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_r1l.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu_r1l.log
python __graft_entry__.py smoke 2>&1 | tail -1
python tools/bench_stages.py hoa_render > gpurun_out/stages_hoa_final.jsonl 2>/dev/null; cut -c1-140 gpurun_out/stages_hoa_final.jsonl
ncu --set full --clock-control none --import-source on -k regex:hoa_render_tc_direct -s 2 -c 1 -o gpurun_out/prof_hoa_render_tc_direct -f \
    python tools/bench_stages.py hoa_render > gpurun_out/ncu_tc_direct.log 2>&1
echo rc=$?
