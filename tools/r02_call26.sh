#!/bin/bash
# GPU call 26: evidence for the final build -- ncu --set full of k_fused_lanes (config 5 at 1/8 batch), ncu launch list of the
# bench command
set -u
out=gpurun_out/r02_call26
mkdir -p $out
: > $out/status.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_fused_lanes" -c 1 \
    -o $out/mlp8_lanes python tools/fullscale_check.py --workload mlp_batch --scale-down 8 --iters 1 --no-check > $out/ncu_mlp8.log 2>&1
echo "ncu mlp8 rc=$?" | tee -a $out/status.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_bench_n1.csv \
    python bench.py --steps 2 --warmup 3 --no-per-config --no-cpu-baseline > $out/ncu_bench.log 2>&1
echo "ncu launches rc=$?" | tee -a $out/status.txt
