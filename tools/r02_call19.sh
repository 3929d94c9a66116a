#!/bin/bash
# GPU call 19: (1) bank-aligned fused dense rows (k_fused_lanes) against the band kernel on config 5, parity + timing + ncu;
# (2) merge chain variants on config 4 / config 2: compare-and-select exchange, 16 keys per lane from class 3 / 4 on
set -u
out=gpurun_out/r02_call19
mkdir -p $out
: > $out/status.txt
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused_lanes or mlp_batch_small" > $out/pytest_fused_lanes.log 2>&1
echo "pytest fused lanes rc=$?" | tee -a $out/status.txt
OSP_FUSED_LANES=1 timeout 600 python tools/fullscale_check.py --workload mlp_batch --iters 4 --sample-rows 12 --heavy-rows 2 --kernels > $out/mlp_batch_lanes.log 2>&1
echo "mlp_batch lanes rc=$?" | tee -a $out/status.txt
OSP_FUSED_LANES=0 timeout 600 python tools/fullscale_check.py --workload mlp_batch --iters 4 --no-check --kernels > $out/mlp_batch_bands.log 2>&1
echo "mlp_batch bands rc=$?" | tee -a $out/status.txt
for lib in default cesel e16from3 e16from4 e16from3_cesel; do
  p=$PWD/gpurun_exp_$lib.so; [ $lib = default ] && p=$PWD/outerspace_b200/libosp_b200.so
  OSP_LIB_PATH=$p timeout 300 python tools/quick_bench.py --workload er8m --iters 5 --flush --kernels --cache > $out/er8m_$lib.log 2>&1
  echo "er8m $lib rc=$?" | tee -a $out/status.txt
  OSP_LIB_PATH=$p timeout 120 python tools/quick_bench.py --workload er16k --iters 8 --flush --kernels --cache --check > $out/er16k_$lib.log 2>&1
  echo "er16k $lib rc=$?" | tee -a $out/status.txt
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_fused_lanes" -c 1 \
    -o $out/mlp8_lanes python tools/fullscale_check.py --workload mlp_batch --scale-down 8 --iters 1 --no-check > $out/ncu_mlp8.log 2>&1
echo "ncu mlp8 rc=$?" | tee -a $out/status.txt
