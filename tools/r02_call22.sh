#!/bin/bash
# GPU call 22: k_fused_lanes with rows of C written at the prefix of their bounds (no look-back) against the chained version;
# three chained sparse MLP layers; the fused-lanes tests
set -u
out=gpurun_out/r02_call22
mkdir -p $out
: > $out/status.txt
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused_lanes or mlp_batch_small or layer_chaining" > $out/pytest_fused_lanes.log 2>&1
echo "pytest fused lanes rc=$?" | tee -a $out/status.txt
timeout 600 python tools/fullscale_check.py --workload mlp_batch --iters 4 --sample-rows 12 --heavy-rows 2 --kernels > $out/mlp_batch_direct.log 2>&1
echo "mlp_batch direct rc=$?" | tee -a $out/status.txt
OSP_FL_DIRECT=0 timeout 600 python tools/fullscale_check.py --workload mlp_batch --iters 4 --sample-rows 4 --heavy-rows 1 --kernels > $out/mlp_batch_lookback.log 2>&1
echo "mlp_batch lookback rc=$?" | tee -a $out/status.txt
timeout 300 python tools/mlp_chain_bench.py > $out/mlp_chain.log 2>&1
echo "mlp chain rc=$?" | tee -a $out/status.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_fused_lanes" -c 1 \
    -o $out/mlp8_lanes python tools/fullscale_check.py --workload mlp_batch --scale-down 8 --iters 1 --no-check > $out/ncu_mlp8.log 2>&1
echo "ncu mlp8 rc=$?" | tee -a $out/status.txt
