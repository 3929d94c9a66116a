#!/bin/bash
# GPU call 9 (8 GPUs): two-phase vs one-phase exchange at N=8, two-phase at N=4
set -u
out=gpurun_out/r02_call9
mkdir -p $out
run() {  # n, one_phase, port
  if [ $2 = 1 ]; then export OSP_DIST_ONE_PHASE=1; else unset OSP_DIST_ONE_PHASE; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $3 bench.py --gpus $1 --steps 10 --warmup 3 \
      > $out/bench_n$1_onephase$2.json 2> $out/bench_n$1_onephase$2.err
  echo "bench n$1 one_phase=$2 rc=$?" | tee -a $out/status.txt
}
run 8 0 29521
run 8 1 29522
run 4 0 29523
