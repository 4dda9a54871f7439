python tools/bench_stages.py > gpurun_out/stages.jsonl 2> gpurun_out/stages.err; echo "stages rc=$? lines=$(wc -l < gpurun_out/stages.jsonl)"
for k in hoa_render_nfc_kernel:2 hoa_render_kernel:2 hoa_render_tc_kernel:40; do
  ncu --set full --clock-control none --import-source on -k regex:${k%%:*} -s ${k##*:} -c 1 -o gpurun_out/prof_${k%%:*} -f \
    python tools/bench_stages.py hoa_render > gpurun_out/ncu_${k%%:*}.log 2>&1
  echo "ncu ${k%%:*} rc=$?"
done
