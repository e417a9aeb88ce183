#!/bin/bash
# ~10 s of box time: the default RK45 build and the experimental 8-warp build (MARLPDE_RK45_BUILD=450) on the
# same 4096-column lattice, each under a hard timeout.   gpurun --timeout 25 -- bash scripts/gpu_quad_contact.sh
set -u
OUT=gpurun_out/quad_contact; mkdir -p $OUT
timeout -s KILL 9 python scripts/gpu_quad_contact.py $OUT/default.npz > $OUT/default.log 2>&1; echo "rc $?"; cat $OUT/default.log
MARLPDE_RK45_BUILD=450 timeout -s KILL 9 python scripts/gpu_quad_contact.py $OUT/quad.npz $OUT/default.npz > $OUT/quad.log 2>&1; echo "rc $?"; cat $OUT/quad.log
