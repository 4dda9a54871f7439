#!/bin/bash
# which part of mpeghb200_fd_execute costs what: MPEGHB200_E2E_DEBUG bit 0 skips the upload, bit 1 the download, bit 2 the kernel
for d in 0 1 2 3 4 7; do
  MPEGHB200_E2E_DEBUG=$d timeout 200 python bench.py --no-cpu-baseline --steps 500 --warmup 10 --e2e-steps 100 2>/dev/null > /tmp/e2e_$d.json
  python -c "import json;d=json.load(open('/tmp/e2e_$d.json'));print($d, d['e2e']['ms_per_step'])"
done
