#!/bin/bash
# round 2, drop-in pass: library-level tests (preload seam + linked drop-in + batch decode)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_dropin_batch.py tests/test_seam_dropin.py -x -q -m gpu -s > gpurun_out/r2b_dropin.log 2>&1; echo "dropin pytest rc=$?"; tail -40 gpurun_out/r2b_dropin.log
