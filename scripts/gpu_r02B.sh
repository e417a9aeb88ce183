#!/bin/bash
# r02B: Radau kernel with global loads that do not allocate in L1 (-Xptxas -dlcm=cg / cs on radau_batch.cu only): the kernel
# keeps ~1 kB of spilled registers per thread in L1 next to 130 GB of streamed vectors and records per launch
set -u
OUT=gpurun_out/${1:-r02B}; mkdir -p $OUT
for i in 1 2; do
  for v in default cg cs; do
    lib=$PWD/build_ab/lib_radau_$v.so; [ $v = default ] && lib=""
    echo "== $v: $(MARLPDE_B200_LIB=$lib timeout 120 python scripts/profile_radau.py 16 0.05 2>&1 | head -1)"
  done
done > $OUT/radau_dlcm.log 2>&1
for v in default cg; do
  lib=$PWD/build_ab/lib_radau_$v.so; [ $v = default ] && lib=""
  echo "== $v, 64 columns: $(MARLPDE_B200_LIB=$lib timeout 120 python scripts/profile_radau.py 4 0.05 2>&1 | head -1)"
done >> $OUT/radau_dlcm.log 2>&1
cat $OUT/radau_dlcm.log
echo done
