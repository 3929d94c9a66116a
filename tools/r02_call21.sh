#!/bin/bash
# GPU call 21: k_fused_lanes v3 (16-byte value loads, predicated quad loads, blocks of 8 groups, rows emitted in place);
# variants: blocks of 4, staged rows; merge chain with four CTAs per SM (64 registers) on config 4
set -u
out=gpurun_out/r02_call21
mkdir -p $out
: > $out/status.txt
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused_lanes or mlp_batch_small or layer_chaining" > $out/pytest_fused_lanes.log 2>&1
echo "pytest fused lanes rc=$?" | tee -a $out/status.txt
timeout 600 python tools/fullscale_check.py --workload mlp_batch --iters 4 --sample-rows 12 --heavy-rows 2 --kernels > $out/mlp_batch_default.log 2>&1
echo "mlp_batch default rc=$?" | tee -a $out/status.txt
for v in fl4 stage; do
  OSP_LIB_PATH=$PWD/gpurun_exp_$v.so timeout 600 python tools/fullscale_check.py --workload mlp_batch --iters 4 --sample-rows 4 --heavy-rows 1 --kernels > $out/mlp_batch_$v.log 2>&1
  echo "mlp_batch $v rc=$?" | tee -a $out/status.txt
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_fused_lanes" -c 1 \
    -o $out/mlp8_lanes python tools/fullscale_check.py --workload mlp_batch --scale-down 8 --iters 1 --no-check > $out/ncu_mlp8.log 2>&1
echo "ncu mlp8 rc=$?" | tee -a $out/status.txt
for lib in occ4 default; do
  p=$PWD/gpurun_exp_$lib.so; [ $lib = default ] && p=$PWD/outerspace_b200/libosp_b200.so
  OSP_LIB_PATH=$p timeout 300 python tools/quick_bench.py --workload er8m --iters 5 --flush --kernels --cache > $out/er8m_$lib.log 2>&1
  echo "er8m $lib rc=$?" | tee -a $out/status.txt
done
