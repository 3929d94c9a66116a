#!/bin/bash
# GPU call 29: the final build once more -- smoke and a parity subset (the narrowed B.pos is now the default from 32 MiB on)
set -u
out=gpurun_out/r02_call29
mkdir -p $out
timeout 45 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "golden or edge_cases or random_vs_oracle or fused_lanes" > $out/pytest_subset.log 2>&1
echo "pytest subset rc=$?" | tee $out/status.txt
