#!/bin/bash
# GPU call 24: final single-GPU record of the round -- whole GPU suite, smoke, the N=1 bench line (per_config with the new
# config-5 path), the ncu launch list of the bench command; blocks of 12 groups for k_fused_lanes
set -u
out=gpurun_out/r02_call24
mkdir -p $out
: > $out/status.txt
OSP_LIB_PATH=$PWD/gpurun_exp_b12.so timeout 600 python tools/fullscale_check.py --workload mlp_batch --iters 5 --sample-rows 4 --heavy-rows 1 --kernels > $out/mlp_batch_b12.log 2>&1
echo "mlp_batch b12 rc=$?" | tee -a $out/status.txt
timeout 1500 python -m pytest tests -m gpu -q > $out/pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a $out/status.txt
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1
echo "smoke rc=$?" | tee -a $out/status.txt
timeout 900 python bench.py --steps 20 --warmup 3 > $out/bench_n1.json 2> $out/bench_n1.err
echo "bench rc=$?" | tee -a $out/status.txt
