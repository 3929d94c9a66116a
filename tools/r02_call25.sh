#!/bin/bash
# GPU call 25 (2 GPUs): the k-sharded path after this session's changes (DevScalars grew, new kernels in the library):
# world-2 GPU tests and the N=2 bench line with per-rank parity
set -u
out=gpurun_out/r02_call25
mkdir -p $out
: > $out/status.txt
timeout 600 python -m pytest tests -m gpu -q -k "world2 or world_2 or sharded or dist" > $out/pytest_world2.log 2>&1
echo "pytest world2 rc=$?" | tee -a $out/status.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 \
    > $out/bench_n2.json 2> $out/bench_n2.err
echo "bench n2 rc=$?" | tee -a $out/status.txt
