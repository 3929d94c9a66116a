#!/bin/bash
# GPU call 5: k_chain2 with asynchronous copies (cp.async) feeding the stages
set -u
out=gpurun_out/r02_call5
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_zzz_fused_short.py -m gpu -x -q > $out/tests.log 2>&1
echo "tests rc=$?" | tee $out/status.txt
if ! grep -q "rc=0" $out/status.txt; then exit 0; fi
for w in "er16k" "er8m --scale-down 8" "er8m"; do
  for f in 0 1; do
    OSP_FUSED_SHORT=$f timeout 300 python tools/quick_bench.py --workload $w --iters 5 --flush --kernels --check \
        > "$out/chain2_${f}_$(echo $w | tr -d ' -').log" 2>&1
    echo "$w fused_short=$f rc=$?" | tee -a $out/status.txt
  done
done
OSP_FUSED_SHORT=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_chain2 -c 1 \
    -o $out/k_chain2_er8m8 python tools/quick_bench.py --workload er8m --scale-down 8 --iters 1 > $out/ncu_chain2.log 2>&1
echo "ncu chain2 rc=$?" | tee -a $out/status.txt
