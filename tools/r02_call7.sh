#!/bin/bash
# GPU call 7 (2 GPUs): whole GPU suite incl. the world-2 test (failure agreement), bench at N=2, k-way merge on config 3 at 1/16
set -u
out=gpurun_out/r02_call7
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee $out/status.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 \
    > $out/bench_n2.json 2> $out/bench_n2.err
echo "bench n2 rc=$?" | tee -a $out/status.txt
for kw in 0 1; do
  OSP_KWAY=$kw timeout 600 python tools/fullscale_check.py --workload rmat20 --scale-down 16 --iters 3 --kernels > $out/rmat16_kway$kw.log 2>&1
  echo "rmat20/16 kway=$kw rc=$?" | tee -a $out/status.txt
done
timeout 300 python tools/quick_bench.py --workload er8m --iters 5 --flush --kernels > $out/er8m_default.log 2>&1
echo "er8m rc=$?" | tee -a $out/status.txt
