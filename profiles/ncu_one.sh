#!/bin/bash
# usage: bash profiles/ncu_one.sh <kernel regex> <tag> [skip]  -> gpurun_out/<tag>.ncu-rep + <tag>_cs.csv (source page) + <tag>_raw.csv
k=$1; tag=$2; skip=${3:-3}
ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/$tag \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_$tag.log 2>&1
ncu -i gpurun_out/$tag.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/${tag}_cs.csv 2>/dev/null
ncu -i gpurun_out/$tag.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
tail -3 gpurun_out/ncu_$tag.log
