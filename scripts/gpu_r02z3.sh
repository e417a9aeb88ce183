#!/bin/bash
# r02z3: is the paired shape slow because its 176 kB of shared memory leave 60 kB of L1 for the spill traffic (classic
# shape: 160 kB -> 92 kB of L1)?  The classic shape with its dynamic shared memory padded into the next carve-out steps.
set -u
OUT=gpurun_out/${1:-r02z3}; mkdir -p $OUT
export MARLPDE_RK45_SHAPE=solo MARLPDE_PROFILE_EVENTS=1
for pad in 0 3000 16000 40000 60000; do
  echo "== classic shape, dynamic shared memory + $pad bytes"
  MARLPDE_RK45_SMEM_PAD=$pad timeout 120 python scripts/profile_rk45.py 3000 2 2>&1 | tail -1
done > $OUT/smem_pad.log 2>&1
cat $OUT/smem_pad.log
echo done
