#!/bin/bash
# GPU call 6: default path after the validate/L2::64B changes; chain geometry experiments (bigger tiles); new bench.py end to end
set -u
out=gpurun_out/r02_call6
mkdir -p $out
for lib in "" gpurun_exp_a.so gpurun_exp_c.so; do
  tag=${lib:-default}
  OSP_LIB_PATH=${lib:+$PWD/$lib} timeout 300 python tools/quick_bench.py --workload er8m --iters 5 --flush --kernels --check > $out/er8m_$tag.log 2>&1
  echo "er8m $tag rc=$?" | tee -a $out/status.txt
done
timeout 300 python tools/quick_bench.py --workload er16k --iters 6 --flush --kernels --check > $out/er16k_default.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > $out/bench_n1.json 2> $out/bench_n1.err
echo "bench rc=$?" | tee -a $out/status.txt
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err
echo "bench ref rc=$?" | tee -a $out/status.txt
