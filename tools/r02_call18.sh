#!/bin/bash
# GPU call 18: symbolic scan with smaller CTAs (fewer warps coupled by the block barriers); e2e with aliased operands
set -u
out=gpurun_out/r02_call18
mkdir -p $out
for b in 32 64 128; do
  OSP_LIB_PATH=$PWD/gpurun_exp_scan$b.so timeout 300 python tools/quick_bench.py --workload er8m --iters 5 --flush --kernels --check > $out/er8m_scan$b.log 2>&1
  echo "er8m scan block $b rc=$?" | tee -a $out/status.txt
done
timeout 600 python bench.py --steps 10 --warmup 3 --no-per-config > $out/bench_n1.json 2> $out/bench_n1.err
echo "bench rc=$?" | tee -a $out/status.txt
