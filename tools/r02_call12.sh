#!/bin/bash
# GPU call 12: how many long-row accumulators should be in flight? (config 3 full size: 4 MB each, L2 = 126 MB)
set -u
out=gpurun_out/r02_call12
mkdir -p $out
for n in 1 2; do
  OSP_XL_CTAS_PER_SM=$n timeout 600 python tools/fullscale_check.py --workload rmat20 --iters 2 --kernels --no-check > $out/rmat20_xl_per_sm$n.log 2>&1
  echo "rmat20 xl per sm $n rc=$?" | tee -a $out/status.txt
done
