#!/bin/bash
# First GPU call of the next round: the opt-in long-row sweep (OSP_LONGROW_SWEEP) has only run on the CPU emulation.
#   gpurun --timeout 1500 -- 'bash tools/r02_first_gpu_call.sh'
# 1. its parity tests (own timeout: a hang must not cost the box);
# 2. config 3 at 1/16 scale and at full size, default path against the sweep, per-kernel event times + sampled rows
#    against the oracle (tools/fullscale_check.py prints them);
# 3. one ncu --set full capture of k_long_fill at 1/16 scale (only if 1 and 2 exited 0).
# Everything lands in gpurun_out/r02_sweep/.
set -u
out=gpurun_out/r02_sweep
mkdir -p $out
OSP_TEST_SWEEP=1 timeout 600 python -m pytest tests/test_gpu_zzz_sweep.py -m gpu -x -q > $out/tests.log 2>&1
echo "sweep tests rc=$?" | tee $out/status.txt
if grep -q "sweep tests rc=0" $out/status.txt; then
for sd in 16 1; do
  for sweep in 0 1; do
    OSP_LONGROW_SWEEP=$sweep timeout 900 python tools/fullscale_check.py --workload rmat20 --scale-down $sd --iters 3 --kernels \
        > $out/rmat20_sd${sd}_sweep${sweep}.log 2>&1
    echo "rmat20/$sd sweep=$sweep rc=$?" | tee -a $out/status.txt
  done
done
# thresholds: every xl row / rows from 32 k / from 256 k partial products
for min in 32768 262144; do
  OSP_LONGROW_SWEEP=1 OSP_LONGROW_SWEEP_MIN=$min timeout 900 python tools/fullscale_check.py --workload rmat20 --scale-down 1 --iters 3 --kernels \
      > $out/rmat20_sd1_sweep1_min${min}.log 2>&1
  echo "rmat20/1 sweep min=$min rc=$?" | tee -a $out/status.txt
done
fi
# the other opt-in path: short-row tiles computed inside the chain (OSP_FUSED_SHORT), configs 2 and 4
OSP_TEST_FUSED_SHORT=1 timeout 600 python -m pytest tests/test_gpu_zzz_fused_short.py -m gpu -x -q > $out/fused_short_tests.log 2>&1
echo "fused short tests rc=$?" | tee -a $out/status.txt
if tail -n 3 $out/fused_short_tests.log | grep -q passed; then
  for w in "er16k" "er8m --scale-down 8" "er8m"; do
    for f in 0 1; do
      OSP_FUSED_SHORT=$f timeout 600 python tools/quick_bench.py --workload $w --iters 4 --flush --kernels --check \
          > "$out/fused${f}_$(echo $w | tr -d ' -').log" 2>&1
      echo "$w fused_short=$f rc=$?" | tee -a $out/status.txt
    done
  done
fi
if ! grep -q "rc=[1-9]" $out/status.txt; then
  OSP_LONGROW_SWEEP=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_long_fill -c 1 \
      -o $out/k_long_fill_rmat16 python tools/fullscale_check.py --workload rmat20 --scale-down 16 --iters 1 --no-check > $out/ncu.log 2>&1
  echo "ncu rc=$?" | tee -a $out/status.txt
fi
