#!/bin/bash
# second A/B round: MARLPDE_RHS_MERGE=3 (two instances of the merged block), warp order, Radau residency / ring depth
set -u
OUT=gpurun_out/${1:-ab_merge2}; mkdir -p $OUT
B=$PWD/build_ab
rk() { MARLPDE_B200_LIB=$B/$1.so MARLPDE_PROFILE_EVENTS=1 timeout 120 python scripts/profile_rk45.py 300 5 > $OUT/rk45_$2.log 2>&1; echo "rk45 $2: $(tail -2 $OUT/rk45_$2.log | tr '\n' ' ')"; }
rd() { MARLPDE_B200_LIB=$B/$1.so timeout 120 python scripts/profile_radau.py 16 0.05 > $OUT/radau_$1.log 2>&1; echo "radau $1: $(head -1 $OUT/radau_$1.log)"; }
rk lib_m0 m0
rk lib_m2 m2
rk lib_m3 m3
MARLPDE_RK45_WARP_PERM=0,1,4,3,2,5,6,7,8,9 rk lib_m3 m3_perm
MARLPDE_RK45_WARP_PERM=0,1,4,3,2,5,6,7,8,9 rk lib_m2 m2_perm
rk lib_m0 m0_again
for l in lib_m2 lib_m3 lib_m2_mb3 lib_m3_mb3 lib_m2_kd2 lib_m2_kd4 lib_m2_mb3_kd4; do rd $l; done
for l in lib_m0 lib_m1 lib_m2 lib_m3; do
  MARLPDE_B200_LIB=$B/$l.so timeout 120 python scripts/profile_stream.py 20000 64 32 > $OUT/tiles_$l.log 2>&1; echo "tiles $l: $(tail -1 $OUT/tiles_$l.log)"
done
MARLPDE_B200_LIB=$B/lib_m3.so timeout 200 python -m pytest -q -m gpu -p no:cacheprovider --timeout=120 --timeout-method=thread tests/test_gpu_rhs.py tests/test_gpu_math.py > $OUT/pytest_lib_m3.log 2>&1
echo "pytest lib_m3: $(tail -1 $OUT/pytest_lib_m3.log)"
echo done
