#!/bin/bash
# stream-level drop-in: 256 handles of the mono and the 12-channel smoke stream through mpeghb200_dec_execute_batch,
# with the scheduler's wall-time split
export LD_LIBRARY_PATH=oracle/_ref:$LD_LIBRARY_PATH MPEGHB200_SEAM_STATS=1
mkdir -p gpurun_out
for s in sine_1khz_000_cicp1.mhas sine_1khz_cicp19.mhas; do
  ./libmpegh_b200/mpeghb200_batch_decode -n:256 -write:0 -o:/tmp/x oracle/_ref/smoke/$s > /tmp/o.json 2> /tmp/e.txt
  python - <<'PY'
import json
d=json.loads(open('/tmp/o.json').read().strip().splitlines()[-1])
print({k:d[k] for k in d if k!='streams'})
PY
  grep "scheduler" /tmp/e.txt
done
