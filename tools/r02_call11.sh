#!/bin/bash
# GPU call 11: ncu --set full of the dominant kernels of configs 4, 3 and 5 (for profiles/traffic.json and the csv summaries),
# launch list of one bench command, then the N=1 bench line of the final build
set -u
out=gpurun_out/r02_call11
mkdir -p $out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_merge_chain|k_multiply|k_scan|k_validate|k_plan" -c 5 \
    -o $out/er8m_default python tools/quick_bench.py --workload er8m --iters 1 > $out/ncu_er8m.log 2>&1
echo "ncu er8m rc=$?" | tee $out/status.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_merge_xl" -c 1 \
    -o $out/rmat16_xl python tools/fullscale_check.py --workload rmat20 --scale-down 16 --iters 1 --no-check > $out/ncu_rmat16.log 2>&1
echo "ncu rmat16 rc=$?" | tee -a $out/status.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_fused_dense" -c 1 \
    -o $out/mlp8_fused python tools/fullscale_check.py --workload mlp_batch --scale-down 8 --iters 1 --no-check > $out/ncu_mlp8.log 2>&1
echo "ncu mlp8 rc=$?" | tee -a $out/status.txt
timeout 900 python bench.py --steps 10 --warmup 3 > $out/bench_n1.json 2> $out/bench_n1.err
echo "bench rc=$?" | tee -a $out/status.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_bench_n1.csv \
    python bench.py --steps 2 --warmup 3 --no-per-config --no-cpu-baseline > $out/ncu_bench.log 2>&1
echo "ncu launches rc=$?" | tee -a $out/status.txt
