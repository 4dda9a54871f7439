ncu --set full --clock-control none --import-source on -k regex:hoa_render_tc_direct -s 2 -c 1 -o gpurun_out/prof_hoa_render_tc_direct -f \
    python tools/bench_stages.py hoa_render > gpurun_out/ncu_tc_direct.log 2>&1
echo rc=$?
