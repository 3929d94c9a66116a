#!/bin/bash
# GPU call 28: symbolic pass with B.pos narrowed to 32 bits (OSP_BPOS32_MIN_KB) on config 4; parity subset with it forced on
set -u
out=gpurun_out/r02_call28
mkdir -p $out
: > $out/status.txt
timeout 85 python tools/quick_bench.py --workload er8m --iters 4 --flush --kernels --cache > $out/er8m_default.log 2>&1
echo "er8m default rc=$?" | tee -a $out/status.txt
OSP_BPOS32_MIN_KB=16384 timeout 40 python tools/quick_bench.py --workload er8m --iters 4 --flush --kernels --cache > $out/er8m_bpos32.log 2>&1
echo "er8m bpos32 rc=$?" | tee -a $out/status.txt
OSP_BPOS32_MIN_KB=1 timeout 40 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "golden or random_vs_oracle or edge_cases or er_config2" > $out/pytest_bpos32.log 2>&1
echo "pytest bpos32 rc=$?" | tee -a $out/status.txt
