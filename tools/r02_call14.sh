#!/bin/bash
# GPU call 14: the sorting network with run-time group size (one body for every row class: code size 134 KB -> 73 KB)
set -u
out=gpurun_out/r02_call14
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zzz_fused_short.py -m gpu -x -q > $out/pytest.log 2>&1
echo "pytest rc=$?" | tee $out/status.txt
for w in "er8m" "er16k" "er8m --scale-down 8"; do
  timeout 300 python tools/quick_bench.py --workload $w --iters 5 --flush --kernels --check > "$out/$(echo $w | tr -d ' -').log" 2>&1
  echo "$w rc=$?" | tee -a $out/status.txt
done
timeout 600 python tools/fullscale_check.py --workload rmat20 --scale-down 16 --iters 3 --kernels > $out/rmat16.log 2>&1
echo "rmat16 rc=$?" | tee -a $out/status.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_merge_chain" -c 1 \
    -o $out/er8m8_chain python tools/quick_bench.py --workload er8m --scale-down 8 --iters 1 > $out/ncu.log 2>&1
echo "ncu rc=$?" | tee -a $out/status.txt
