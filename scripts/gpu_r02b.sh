#!/bin/bash
# r02b: (1) 128-thread CTAs, one column each, 3 per SM (MARLPDE_RK45_BUILD=128; 129 = y, K1 in shared memory) against the
# 320-thread / 3-column CTA; (2) quantum-major work items (MARLPDE_FLAG_QUEUE_LOCKS) against whole-column claims.
set -u
OUT=gpurun_out/${1:-r02b}; mkdir -p $OUT
run() { # build quantum attempts launches lattice tag
  MARLPDE_RK45_BUILD=$1 MARLPDE_PROFILE_QUANTUM=$2 MARLPDE_PROFILE_LATTICE=$5 MARLPDE_PROFILE_EVENTS=1 timeout 150 python scripts/profile_rk45.py $3 $4 > $OUT/$6.log 2>&1
  echo "$6: $(tail -2 $OUT/$6.log | tr '\n' ' ')"
}
for b in 320 128 129 321; do
  run $b -1 300 5 16,16,16 b${b}_300
done
for b in 320 128; do
  for q in -1 0 250; do
    run $b $q 3000 3 16,16,16 b${b}_q${q}_3000
  done
  run $b -1 3000 3 4,4,4 b${b}_64col
  run $b -1 3000 3 16,16,32 b${b}_8192col_qoff
  run $b 0 3000 3 16,16,32 b${b}_8192col_qauto
done
MARLPDE_RK45_BUILD=128 timeout 600 python -m pytest -q -x -m gpu -p no:cacheprovider --timeout=120 --timeout-method=thread tests/test_gpu_rk45.py tests/test_gpu_dropin.py > $OUT/pytest_128.log 2>&1
echo "pytest 128: $(tail -1 $OUT/pytest_128.log)"
timeout 600 python -m pytest -q -x -m gpu -p no:cacheprovider --timeout=120 --timeout-method=thread tests/test_gpu_rk45.py tests/test_gpu_dropin.py tests/test_gpu_reference_suite.py > $OUT/pytest_320.log 2>&1
echo "pytest 320: $(tail -1 $OUT/pytest_320.log)"
MARLPDE_RK45_BUILD=128 MARLPDE_PROFILE_EVENTS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:rk45_persistent -s 1 -c 1 -o $OUT/rk45_128_full python scripts/profile_rk45.py 300 3 > $OUT/ncu_128.log 2>&1; echo "ncu 128: rc $?"
echo done
