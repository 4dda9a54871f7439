#!/bin/bash
# compute-sanitizer over the GPU parity tests: memcheck on the whole chain / stage suites that
# finish in reasonable time under the tool, racecheck on the kernels that use named barriers, cp.async rings, PDL,
# mbarriers.  Summaries go to gpurun_out/ (copied to profiles/ by hand).
mkdir -p gpurun_out
SUBSET="tests/test_chain_gpu.py tests/test_tail_gpu.py tests/test_fd_synth_gpu.py tests/test_obj_render_gpu.py tests/test_hoa_render_gpu.py tests/test_mix_gpu.py tests/test_ltpf_gpu.py"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 97 --log-file gpurun_out/r2_memcheck.log \
  python -m pytest $SUBSET -x -q -m gpu -k "not full_size and not config2_shape and not 256" > gpurun_out/r2_memcheck_pytest.log 2>&1
echo "memcheck rc=$?"; tail -2 gpurun_out/r2_memcheck_pytest.log; grep -E "ERROR SUMMARY|Invalid|Misaligned" gpurun_out/r2_memcheck.log | sort | uniq -c | head
RSUB="tests/test_fd_synth_gpu.py tests/test_hoa_render_gpu.py tests/test_tail_gpu.py tests/test_obj_render_gpu.py"
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 97 --log-file gpurun_out/r2_racecheck.log \
  python -m pytest $RSUB -x -q -m gpu -k "golden or umma_mode_vs_oracle or fused or streams_vs_oracle" > gpurun_out/r2_racecheck_pytest.log 2>&1
echo "racecheck rc=$?"; tail -2 gpurun_out/r2_racecheck_pytest.log; grep -E "RACECHECK SUMMARY|hazard" gpurun_out/r2_racecheck.log | sort | uniq -c | head
