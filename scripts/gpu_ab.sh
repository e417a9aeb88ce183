#!/bin/bash
set -u
OUT=gpurun_out/${1:-ab}
mkdir -p $OUT
for v in 000 100 010 001; do
  MARLPDE_B200_LIB=$PWD/build_variants/lib_$v.so timeout 300 python scripts/profile_radau.py 16 0.05 > $OUT/v$v.log 2>&1
done
echo done
