#!/bin/bash
# GPU call 8 (2 GPUs): two-phase exchange (owners merge the first half while the second is sent)
set -u
out=gpurun_out/r02_call8
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > $out/pytest_dist.log 2>&1
echo "pytest dist rc=$?" | tee $out/status.txt
for one in 0 1; do
  if [ $one = 1 ]; then export OSP_DIST_ONE_PHASE=1; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$one bench.py --gpus 2 --steps 10 --warmup 3 \
      > $out/bench_n2_onephase$one.json 2> $out/bench_n2_onephase$one.err
  echo "bench n2 one_phase=$one rc=$?" | tee -a $out/status.txt
done
