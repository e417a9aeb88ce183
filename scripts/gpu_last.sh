#!/bin/bash
# last call of r01: twisted Radau residency A/B, full GPU parity suite on the in-tree build, implicit sweep to T*
set -u
OUT=gpurun_out/${1:-last}; mkdir -p $OUT
for v in tw_mb2 tw_mb3; do
  MARLPDE_B200_LIB=$PWD/build_ab/lib_$v.so timeout 100 python scripts/profile_radau.py 16 0.05 > $OUT/radau_$v.log 2>&1; echo "radau $v: $(head -1 $OUT/radau_$v.log)"
done
timeout 100 python scripts/profile_radau.py 16 0.05 > $OUT/radau_intree.log 2>&1; echo "radau in-tree (mb4): $(head -1 $OUT/radau_intree.log)"
( time timeout 600 python -m pytest -q -m gpu -p no:cacheprovider --timeout=300 --timeout-method=thread tests ) > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu.log; tail -5 $OUT/pytest_gpu.log
timeout 150 python scripts/profile_radau.py 16 1.0 > $OUT/radau_tstar.log 2>&1; echo "implicit T* default base: $(head -1 $OUT/radau_tstar.log)"
echo done
