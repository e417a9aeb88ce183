#!/bin/bash
# third A/B round: kSchedLean (voted weight selection), integer range check in exp_nb
set -u
OUT=gpurun_out/${1:-ab_merge3}; mkdir -p $OUT
B=$PWD/build_ab
rk() { MARLPDE_B200_LIB=$B/$1.so MARLPDE_PROFILE_EVENTS=1 timeout 120 python scripts/profile_rk45.py 300 5 > $OUT/rk45_$2.log 2>&1; echo "rk45 $2: $(tail -2 $OUT/rk45_$2.log | tr '\n' ' ')"; }
rd() { MARLPDE_B200_LIB=$B/$1.so timeout 120 python scripts/profile_radau.py 16 0.05 > $OUT/radau_$1.log 2>&1; echo "radau $1: $(head -1 $OUT/radau_$1.log)"; }
rk lib_m3 m3
rk lib_m4 m4
rk lib_m4i m4i
rk lib_m3i m3i
MARLPDE_RK45_WARP_PERM=0,1,4,3,2,5,6,7,8,9 rk lib_m4 m4_perm
rk lib_m3 m3_again
rd lib_m4
rd lib_m2
MARLPDE_B200_LIB=$B/lib_m4i.so timeout 200 python -m pytest -q -m gpu -p no:cacheprovider --timeout=120 --timeout-method=thread tests/test_gpu_rhs.py tests/test_gpu_math.py > $OUT/pytest_lib_m4i.log 2>&1
echo "pytest lib_m4i: $(tail -1 $OUT/pytest_lib_m4i.log)"
echo done
