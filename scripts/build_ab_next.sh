#!/bin/bash
# Variant libraries for scripts/gpu_ab_next.sh (build_ab/ is git-ignored but travels to the GPU box).
set -eu
cd "$(dirname "$0")/../integrating-diagenetic-equations-using-python_b200"
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -cudart static --threads 0"
mkdir -p ../build_ab
$NV -DMARLPDE_FP64_IMM=1 -o ../build_ab/lib_imm.so csrc/*.cu &
$NV -DMARLPDE_RADAU_FUSE_F=1 -o ../build_ab/lib_fuse.so csrc/*.cu &
$NV -DMARLPDE_TAIL_SPREAD=1 -o ../build_ab/lib_spread.so csrc/*.cu &
$NV -DMARLPDE_TILE_TMA=1 -o ../build_ab/lib_tma.so csrc/*.cu &
$NV -DMARLPDE_QUAD_ROLLED=1 -o ../build_ab/lib_quad_rolled.so csrc/*.cu &
$NV -DMARLPDE_QUAD_ORDER=1 -o ../build_ab/lib_quad_o1.so csrc/*.cu &
for m in 2 3; do $NV -DMARLPDE_RHS_MERGE=$m -o ../build_ab/lib_m$m.so csrc/*.cu & done
wait
