python -m pytest tests/test_hoa_render_gpu.py -x -q -k "not tensor" 2>&1 | tail -5
python tools/bench_stages.py hoa_render 2>&1 | cut -c1-200
