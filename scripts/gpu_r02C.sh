#!/bin/bash
# r02C: Radau kernel without the cp.async staging of the RHS inputs AND without its 10.7 kB per warp of shared memory:
# 41 kB instead of 47 kB per CTA moves the SM (3 CTAs) from the 164 kB to the 132 kB carve-out: 124 kB of L1 instead of 92 kB
set -u
OUT=gpurun_out/${1:-r02C}; mkdir -p $OUT
for i in 1 2; do
  for v in default nostage; do
    lib=$PWD/build_ab/lib_implicit_$v.so; [ $v = default ] && lib=""
    echo "== $v: $(MARLPDE_B200_LIB=$lib timeout 120 python scripts/profile_radau.py 16 0.05 2>&1 | head -1)"
  done
done > $OUT/radau_l1.log 2>&1
for v in default nostage; do
  lib=$PWD/build_ab/lib_implicit_$v.so; [ $v = default ] && lib=""
  echo "== $v, 64 columns: $(MARLPDE_B200_LIB=$lib timeout 120 python scripts/profile_radau.py 4 0.05 2>&1 | head -1)"
done >> $OUT/radau_l1.log 2>&1
cat $OUT/radau_l1.log
echo done
