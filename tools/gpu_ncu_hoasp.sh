mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'hoasp_render' -s 20 -c 1 -o gpurun_out/prof_hoasp2 -f \
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 3 --chain-steps 12 --chain-e2e-steps 3 --only hoa > gpurun_out/ncu_hoasp2.log 2>&1
echo rc=$?
