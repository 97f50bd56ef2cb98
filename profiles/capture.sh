#!/bin/bash
# Reproduces the round's evidence on one B200 (run from the repo root on the GPU box; outputs under gpurun_out/).
set -u
O=gpurun_out; mkdir -p $O
python bench.py > $O/r1_bench_1gpu.json 2> $O/r1_bench_1gpu.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/r1_bench_reference.json 2>> $O/r1_bench_1gpu.err
for w in c2 c3 c4; do python bench.py --workload $w --steps 20 > $O/r1_bench_${w}_1gpu.json 2>> $O/r1_bench_1gpu.err; done
python bench.py --workload c5 --steps 5 > $O/r1_bench_c5_1gpu.json 2>> $O/r1_bench_1gpu.err
python bench.py --workload bow --steps 20 > $O/r1_bench_bow_1gpu.json 2>> $O/r1_bench_1gpu.err
python profiles/h2d_probe.py > $O/r1_h2d_probe.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r1_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
for k in k_pyr_fast k_fast_seg k_octree k_orient_describe k_bf_knn2 k_window_candidates k_window_resolve; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o $O/r1_$k \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $O/ncu_$k.log 2>&1
done
# the level-0 blur-only instantiation of the pyramid kernel is the first launch of a step
ncu --set full --clock-control none --import-source on -k regex:k_pyr_fast -s 24 -c 1 -f -o $O/r1_k_pyr_fast_l0 \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $O/ncu_k_pyr_fast_l0.log 2>&1
for k in k_fast_seg k_pyr_fast; do
  ncu -i $O/r1_$k.ncu-rep --page source --print-source cuda,sass --csv > $O/r1_${k}_cs.csv 2>/dev/null
done
ls -la $O | tail -30
