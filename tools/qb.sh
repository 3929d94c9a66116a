#!/bin/bash
# usage: tools/qb.sh TAG "workload args" ...   -> gpurun_out/qb_TAG_*.log (iteration 2 + kernel table of the last)
tag=$1; shift
for w in "$@"; do
  f="gpurun_out/qb_${tag}_$(echo $w | tr -d ' -').log"
  timeout 600 python tools/quick_bench.py --workload $w --iters 4 --flush --kernels --check > "$f" 2>&1
  echo "== $w"
  grep -E "oracle|Error|error" "$f" | head -3
  grep '"it": 2' "$f" | python -c "import sys,json; [print({k:d[k] for k in ('gflops','alg_gbs','ms_total','ms_convert','ms_multiply','ms_merge','merge_tiles','kernel_launches')}) for d in map(json.loads, sys.stdin)]"
  grep " us " "$f"
done
