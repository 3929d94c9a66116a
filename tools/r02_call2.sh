#!/bin/bash
# GPU call 2 of round 2: gather probe (+ncu bytes), full GPU suite with the opt-in paths un-gated, L2 fetch granularity on er8m.
set -u
out=gpurun_out/r02_call2
mkdir -p $out
timeout 600 tools/probe_gather.bin 33554432 > $out/probe_gather.log 2>&1
echo "probe rc=$?" | tee $out/status.txt
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 70 --csv \
    --log-file $out/probe_gather_ncu.csv tools/probe_gather.bin 8388608 > $out/probe_gather_ncu.log 2>&1
echo "probe ncu rc=$?" | tee -a $out/status.txt
timeout 1500 python -m pytest tests -m gpu -q -x > $out/pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a $out/status.txt
for g in 0 32 128; do
  OSP_L2_FETCH=$g timeout 600 python tools/quick_bench.py --workload er8m --iters 4 --flush --kernels > $out/er8m_l2fetch$g.log 2>&1
  echo "er8m l2fetch=$g rc=$?" | tee -a $out/status.txt
done
