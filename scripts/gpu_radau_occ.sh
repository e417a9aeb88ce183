#!/bin/bash
set -u
OUT=gpurun_out/${1:-radau_occ}
mkdir -p $OUT
for mb in 4 6; do
  MARLPDE_B200_LIB=$PWD/build_variants/libmarlpde_radau_mb$mb.so timeout 300 python scripts/profile_radau.py 16 0.05 > $OUT/mb$mb.log 2>&1
done
timeout 300 python scripts/profile_radau.py 16 0.05 > $OUT/mb8.log 2>&1
timeout 300 python scripts/profile_radau.py 4 0.05 > $OUT/mb8_64.log 2>&1
echo done
