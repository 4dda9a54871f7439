#!/bin/bash
# Stream-level throughput of the drop-in: N handles of a smoke stream through mpeghb200_dec_execute_batch (parse in the
# reference's host code on the parse pool, signal path on the GPU), next to the unmodified testbench on the host cores.
mkdir -p gpurun_out
export LD_LIBRARY_PATH=$PWD/oracle/_ref
OUT=gpurun_out/r2e_stream_level.jsonl; : > $OUT
for s in sine_1khz_000_cicp1.mhas sine_1khz_cicp19.mhas; do
  for n in 1 64 256 1024; do
    for t in 1 16; do
      if [ $t = 1 ] && [ $n != 256 ]; then continue; fi
      MPEGHB200_PARSE_THREADS=$t libmpegh_b200/mpeghb200_batch_decode -n:$n -write:0 -o:/tmp/sl oracle/_ref/smoke/$s 2>/dev/null |
        python -c "import sys,json; l=json.loads(sys.stdin.read()); print(json.dumps({'stream':'$s','handles':l['handles'],'parse_threads':$t,'frames':l['frames'],'run_s':l['run_s'],'decode_s':l['decode_s'],'audio_s_per_decode_s':l['audio_s_per_decode_s'],'init_s':l['init_s'],'audio_s_per_wall_s':l['audio_s_per_wall_s'],'channels':l['streams'][0]['channels']}))" | tee -a $OUT
    done
  done
done
python - <<'PY' | tee -a gpurun_out/r2e_stream_level.jsonl
import sys, json
sys.path.insert(0, '.')
sys.argv = ['x']
import bench
print(json.dumps({"cpu_testbench": bench.cpu_testbench_baseline(bench.cpu_threads())}))
print(json.dumps({"cpu_testbench": bench.cpu_testbench_baseline(bench.cpu_threads(), 4, "sine_1khz_cicp19.mhas", 12)}))
PY
