#!/bin/bash
# new factorise: parity tests, then timing of build variants on the same box (t = 0.05, 4096 columns, default base)
set -u
OUT=gpurun_out/${1:-radau_fact}; mkdir -p $OUT
timeout 600 python -m pytest -q -m gpu -p no:cacheprovider --timeout=300 --timeout-method=thread tests/test_gpu_radau.py tests/test_gpu_dropin.py > $OUT/pytest.log 2>&1
echo "pytest exit $?" >> $OUT/pytest.log
timeout 120 python scripts/profile_radau.py 16 0.05 > $OUT/new.log 2>&1
for v in old pfrhs pfsolve mb5; do
  f=$PWD/build_ab/lib_$v.so; [ $v = old ] && f=$PWD/build_ab/libold.so
  MARLPDE_B200_LIB=$f timeout 120 python scripts/profile_radau.py 16 0.05 > $OUT/$v.log 2>&1
done
timeout 120 python scripts/profile_radau.py 4 0.05 > $OUT/new_64.log 2>&1
grep -h "columns to" $OUT/*.log; tail -3 $OUT/pytest.log
