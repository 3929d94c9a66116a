#!/bin/bash
# GPU call 3: k_chain2 (warp-specialised fused chain): parity, timing against the default path, ncu capture
set -u
out=gpurun_out/r02_call3
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_zzz_fused_short.py -m gpu -x -q > $out/fused_short_tests.log 2>&1
echo "chain2 tests rc=$?" | tee $out/status.txt
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
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_merge_chain|k_multiply|k_scan" -c 3 \
    -o $out/default_er8m8 python tools/quick_bench.py --workload er8m --scale-down 8 --iters 1 > $out/ncu_default.log 2>&1
echo "ncu default rc=$?" | tee -a $out/status.txt
