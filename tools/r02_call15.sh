#!/bin/bash
# GPU call 15: scan with 16 items per thread (experiment), then the whole GPU suite and the N=1 bench line of the final build
set -u
out=gpurun_out/r02_call15
mkdir -p $out
for lib in "" gpurun_exp_f.so; do
  tag=${lib:-default}
  OSP_LIB_PATH=${lib:+$PWD/$lib} timeout 300 python tools/quick_bench.py --workload er8m --iters 5 --flush --kernels --check > $out/er8m_$tag.log 2>&1
  echo "er8m $tag rc=$?" | tee -a $out/status.txt
done
timeout 1500 python -m pytest tests -m gpu -q > $out/pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a $out/status.txt
timeout 900 python bench.py --steps 20 --warmup 3 > $out/bench_n1.json 2> $out/bench_n1.err
echo "bench rc=$?" | tee -a $out/status.txt
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1
echo "smoke rc=$?" | tee -a $out/status.txt
