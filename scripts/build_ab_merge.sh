#!/bin/bash
# Variant libraries for scripts/gpu_ab_merge.sh (build_ab/ is git-ignored but travels to the GPU box).
set -eu
cd "$(dirname "$0")/../integrating-diagenetic-equations-using-python_b200"
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -cudart static --threads 0"
mkdir -p ../build_ab
for m in 0 1 2; do $NV -DMARLPDE_RHS_MERGE=$m -o ../build_ab/lib_m$m.so csrc/*.cu & done
for mb in 3 5; do $NV -DMARLPDE_RHS_MERGE=0 -DMARLPDE_RADAU_MINBLOCKS=$mb -o ../build_ab/lib_m0_mb$mb.so csrc/*.cu & done
$NV -DMARLPDE_RHS_MERGE=2 -DMARLPDE_RADAU_MINBLOCKS=5 -o ../build_ab/lib_m2_mb5.so csrc/*.cu &
wait
