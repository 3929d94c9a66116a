#!/bin/bash
# GPU call 23: load pipeline of k_fused_lanes: two or three register sets, 8 or 6 quads per set (config 5 at full size)
set -u
out=gpurun_out/r02_call23
mkdir -p $out
: > $out/status.txt
for v in default d3q6 d3q8 d2q6; do
  p=$PWD/gpurun_exp_$v.so; [ $v = default ] && p=$PWD/outerspace_b200/libosp_b200.so
  OSP_LIB_PATH=$p timeout 600 python tools/fullscale_check.py --workload mlp_batch --iters 5 --sample-rows 4 --heavy-rows 1 --kernels > $out/mlp_batch_$v.log 2>&1
  echo "mlp_batch $v rc=$?" | tee -a $out/status.txt
done
