#!/bin/bash
# usage: ncu_stage.sh <bench_stages.py stage> <kernel regex> <skip> <tag>  -- one full ncu capture of a stage kernel
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$2 -s ${3:-2} -c 1 -o gpurun_out/prof_$4 -f \
    python tools/bench_stages.py $1 > gpurun_out/ncu_$4.log 2>&1
echo "ncu $4 rc=$?"
