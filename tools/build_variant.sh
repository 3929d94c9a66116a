#!/bin/bash
# Builds an experimental copy of the engine library with extra -D flags: tools/build_variant.sh <name> [-DFLAG=V ...]
# -> gpurun_exp_<name>.so at the repo root (git-ignored, travels to the GPU box; select it with OSP_LIB_PATH).
set -e
name=$1; shift
cd "$(dirname "$0")/../outerspace_b200/csrc"
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-O3,-Wall -Xptxas -v --fmad=false "$@" \
    -shared -o ../../gpurun_exp_$name.so osp_engine.cu osp_host.cpp -lcudart 2> /tmp/ptxas_$name.log || (cat /tmp/ptxas_$name.log; exit 1)
grep -A2 "k_merge_chainIjLb0" /tmp/ptxas_$name.log | grep -E "Used|spill" | tr '\n' ' '; echo " [$name]"
