#!/bin/bash
# GPU call 17 (2 GPUs): final build -- whole GPU suite incl. world 2 (two-phase exchange forced), er8m kernels, bench at N=2 and N=1
set -u
out=gpurun_out/r02_call17
mkdir -p $out
timeout 300 python tools/quick_bench.py --workload er8m --iters 5 --flush --kernels --check > $out/er8m.log 2>&1
echo "er8m rc=$?" | tee $out/status.txt
timeout 1500 python -m pytest tests -m gpu -q > $out/pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a $out/status.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 \
    > $out/bench_n2.json 2> $out/bench_n2.err
echo "bench n2 rc=$?" | tee -a $out/status.txt
timeout 600 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err
echo "bench ref rc=$?" | tee -a $out/status.txt
