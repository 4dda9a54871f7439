#!/bin/bash
# Stream-level drop-in on G GPUs of one box: one process per GPU (CUDA_VISIBLE_DEVICES), 256 handles each through
# mpeghb200_dec_execute_batch, the host's cores split between the processes' parse pools.  Aggregate = audio-seconds of
# all processes / the slowest process' decode time (they start together).
mkdir -p gpurun_out
export LD_LIBRARY_PATH=$PWD/oracle/_ref
OUT=gpurun_out/r2e_stream_level_multi.jsonl; : > $OUT
NCPU=$(nproc)
NGPU=$(nvidia-smi -L | wc -l)
for s in sine_1khz_000_cicp1.mhas sine_1khz_cicp19.mhas; do
  for G in 1 2 4 8; do
    [ $G -gt $NGPU ] && continue
    T=$(( NCPU / G )); [ $T -gt 16 ] && T=16; [ $T -lt 1 ] && T=1
    for g in $(seq 0 $((G-1))); do
      CUDA_VISIBLE_DEVICES=$g MPEGHB200_PARSE_THREADS=$T libmpegh_b200/mpeghb200_batch_decode -n:256 -write:0 -o:/tmp/slm$g \
        oracle/_ref/smoke/$s > /tmp/slm_$g.json 2>/dev/null &
    done
    wait
    python - "$s" $G $T <<'PY' | tee -a $OUT
import json, sys
s, G, T = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
rs = [json.loads(open(f"/tmp/slm_{g}.json").read().strip().splitlines()[-1]) for g in range(G)]
audio = sum(r["frames"] for r in rs) * 1024 / 48000.0
print(json.dumps({"stream": s, "gpus": G, "handles_per_gpu": 256, "parse_threads_per_gpu": T,
                  "audio_s_per_decode_s": audio / max(r["decode_s"] for r in rs),
                  "audio_s_per_wall_s": audio / max(r["run_s"] for r in rs),
                  "decode_s": [r["decode_s"] for r in rs]}))
PY
  done
done
