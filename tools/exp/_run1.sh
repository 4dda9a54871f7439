python -m pytest tests/test_hoa_render_gpu.py -x -q -s 2>&1 | grep -v "^$" | tail -12
python tools/bench_stages.py hoa_render 2>&1 | cut -c1-200
