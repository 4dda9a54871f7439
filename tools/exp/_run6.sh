python -m pytest tests/test_obj_render_gpu.py -x -q 2>&1 | tail -3
python tools/bench_stages.py obj 2>&1 | cut -c1-220
