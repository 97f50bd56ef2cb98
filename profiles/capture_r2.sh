#!/bin/bash
# Reproduces the round-2 evidence on one B200 (run from the repo root on the GPU box; outputs under gpurun_out/).
set -u
O=gpurun_out; mkdir -p $O
python bench.py > $O/r2_bench_1gpu.json 2> $O/r2_bench_1gpu.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_bench_reference.json 2>> $O/r2_bench_1gpu.err
for w in c2 c3 c4; do python bench.py --workload $w --steps 20 > $O/r2_bench_${w}_1gpu.json 2>> $O/r2_bench_1gpu.err; done
python bench.py --workload c5 --steps 5 > $O/r2_bench_c5_1gpu.json 2>> $O/r2_bench_1gpu.err
python bench.py --workload bow --steps 20 > $O/r2_bench_bow_1gpu.json 2>> $O/r2_bench_1gpu.err
PYTHONPATH=. python profiles/latency_probe.py 1000 > $O/r2_latency.txt 2>&1
ORBX_DAG=0 PYTHONPATH=. python profiles/latency_probe.py 1000 | sed 's/^/serial launch order inside the graph (ORBX_DAG=0): /' >> $O/r2_latency.txt 2>&1
ORBX_NO_GRAPH=1 PYTHONPATH=. python profiles/latency_probe.py 1000 | sed 's/^/direct launches (ORBX_NO_GRAPH=1): /' >> $O/r2_latency.txt 2>&1
PYTHONPATH=. python profiles/latency_stages.py 500 >> $O/r2_latency.txt 2>&1
python profiles/e2e_probe.py > $O/r2_e2e_probe.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r2_latency_launches.csv \
    python profiles/latency_probe.py 30 > /dev/null 2>&1
for k in k_pyr_fast k_fast_seg k_octree k_orient k_describe k_bf_knn2_tcws k_window_candidates k_init_resolve; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o $O/r2_$k \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $O/ncu_$k.log 2>&1
done
# the level-0 blur-only instantiation of the pyramid kernel is the first launch of a step
ncu --set full --clock-control none --import-source on -k regex:k_pyr_fast -s 24 -c 1 -f -o $O/r2_k_pyr_fast_l0 \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $O/ncu_k_pyr_fast_l0.log 2>&1
# the tensor-core kernel at the C5 scale (one launch = 1000 queries x 65.5 M descriptors)
ncu --set full --clock-control none --import-source on -k regex:k_bf_knn2_tcws -s 3 -c 1 -f -o $O/r2_k_bf_knn2_tc_c5 \
    python bench.py --workload c5 --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_bftc_c5.log 2>&1
for k in k_fast_seg k_pyr_fast k_bf_knn2_tc_c5; do
  ncu -i $O/r2_$k.ncu-rep --page source --print-source cuda,sass --csv > $O/r2_${k}_cs.csv 2>/dev/null
done
python profiles/ncu_summary.py $O/r2_ncu_summary.json $O/r2_k_pyr_fast.ncu-rep $O/r2_k_pyr_fast_l0.ncu-rep $O/r2_k_fast_seg.ncu-rep $O/r2_k_octree.ncu-rep \
    $O/r2_k_orient.ncu-rep $O/r2_k_describe.ncu-rep $O/r2_k_bf_knn2_tcws.ncu-rep $O/r2_k_bf_knn2_tc_c5.ncu-rep $O/r2_k_window_candidates.ncu-rep \
    $O/r2_k_init_resolve.ncu-rep > $O/r2_ncu_summary.txt 2>&1
python profiles/ncu_lines.py $O/r2_k_fast_seg_cs.csv 45 > $O/r2_fast_lines.txt 2>&1
python profiles/ncu_lines.py $O/r2_k_pyr_fast_cs.csv 45 > $O/r2_pyr_lines.txt 2>&1
python profiles/ncu_lines.py $O/r2_k_bf_knn2_tc_c5_cs.csv 30 > $O/r2_bftc_lines.txt 2>&1
python profiles/summarize_launches.py $O/r2_launches.csv > $O/r2_launches_summary.txt 2>&1
python profiles/summarize_launches.py $O/r2_latency_launches.csv > $O/r2_latency_launches_summary.txt 2>&1
# memcheck over the kernels added this round (the tensor-core kNN, the rig searches, the keyframe DB) and the extraction suite
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_match_gpu.py tests/test_kfdb_gpu.py tests/test_extract_gpu.py tests/test_server_gpu.py -m gpu -x -q \
    -k "knn or two_camera or kfdb or prefetch or projection or parity or pipeline or partial or graph" > $O/r2_sanitizer.txt 2>&1; echo "memcheck exit code $?" >> $O/r2_sanitizer.txt
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_match_gpu.py -m gpu -x -q -k "knn or two_camera or host_pipeline or graph" > $O/r2_racecheck.txt 2>&1
echo "racecheck exit code $?" >> $O/r2_racecheck.txt
# the level-parallel launch order forced onto direct launches (per-level FAST / octree / blur on branch streams), under memcheck
ORBX_DAG=1 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_extract_gpu.py tests/test_match_gpu.py -m gpu -x -q \
    -k "graph or host_pipeline or single or parity" > $O/r2_sanitizer_dag.txt 2>&1; echo "memcheck (ORBX_DAG=1) exit code $?" >> $O/r2_sanitizer_dag.txt
# the host pipeline's kNN tables against the oracle under the sanitizer's timing (the run that exposed the key-base race)
timeout 300 compute-sanitizer --tool memcheck python profiles/pipeline_knn_check.py > $O/r2_pipeline_knn_check.txt 2>&1
ls -la $O | tail -40
