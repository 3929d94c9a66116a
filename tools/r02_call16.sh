#!/bin/bash
# GPU call 16 (4 GPUs): two-phase exchange with the merge chain of the overlapped half capped at 2 CTAs per SM vs 3
set -u
out=gpurun_out/r02_call16
mkdir -p $out
i=0
for occ in 3 2 1; do
  i=$((i+1))
  OSP_DIST_CHAIN_OCC=$occ timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 2953$i bench.py --gpus 4 --steps 10 --warmup 3 \
      > $out/bench_n4_occ$occ.json 2> $out/bench_n4_occ$occ.err
  echo "bench n4 occ=$occ rc=$?" | tee -a $out/status.txt
done
