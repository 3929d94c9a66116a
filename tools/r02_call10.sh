#!/bin/bash
# GPU call 10: default path after the scan rewrite and the early look-back resolve; occupancy experiments; GPU suite
set -u
out=gpurun_out/r02_call10
mkdir -p $out
for lib in "" gpurun_exp_d.so gpurun_exp_e.so; do
  tag=${lib:-default}
  OSP_LIB_PATH=${lib:+$PWD/$lib} timeout 300 python tools/quick_bench.py --workload er8m --iters 5 --flush --kernels --check > $out/er8m_$tag.log 2>&1
  echo "er8m $tag rc=$?" | tee -a $out/status.txt
  OSP_LIB_PATH=${lib:+$PWD/$lib} timeout 300 python tools/quick_bench.py --workload er16k --iters 6 --flush --kernels --check > $out/er16k_$tag.log 2>&1
done
timeout 1500 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a $out/status.txt
