#!/bin/bash
# r02F: on-chip RK45 kernel with the thread's position packed into one opaque word (slot / pair index / first / last from bit
# operations instead of the re-derived division chains) and has1 compiled out for even N, against the previous build
# (build_ab/lib_head.so): the bench step (4096 columns x 3000 attempts, events on), then the RK45 parity tests
set -u
OUT=gpurun_out/${1:-r02F}; mkdir -p $OUT
export MARLPDE_PROFILE_EVENTS=1
for i in 1 2; do
  echo "== new:  $(timeout 120 python scripts/profile_rk45.py 3000 2 2>&1 | tail -1)"
  echo "== head: $(MARLPDE_B200_LIB=$PWD/build_ab/lib_head.so timeout 120 python scripts/profile_rk45.py 3000 2 2>&1 | tail -1)"
done > $OUT/where_ab.log 2>&1
cat $OUT/where_ab.log
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout=400 --timeout-method=thread"
( time timeout 600 $PT tests/test_gpu_rk45.py ) > $OUT/pytest_rk45.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_rk45.log; tail -5 $OUT/pytest_rk45.log
echo done
