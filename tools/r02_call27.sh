#!/bin/bash
# GPU call 27: compute-sanitizer (memcheck, then racecheck) over the bank-aligned fused dense rows on small cases
set -u
out=gpurun_out/r02_call27
mkdir -p $out
: > $out/status.txt
timeout 75 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused_lanes_bank_aligned_rows and (4096 or 999 or 8160)" > $out/memcheck.log 2>&1
echo "memcheck rc=$?" | tee -a $out/status.txt
timeout 55 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused_lanes_bank_aligned_rows and 4096-1-1" > $out/racecheck.log 2>&1
echo "racecheck rc=$?" | tee -a $out/status.txt
