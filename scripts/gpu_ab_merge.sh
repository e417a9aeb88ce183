#!/bin/bash
# A/B of the RHS scheduling variants (MARLPDE_RHS_MERGE 0/1/2), the warp order of the RK45 kernel and the Radau
# residency, all as separate libraries under build_ab/ (built by scripts/build_ab_merge.sh).
set -u
OUT=gpurun_out/${1:-ab_merge}; mkdir -p $OUT
B=$PWD/build_ab
rk() { MARLPDE_B200_LIB=$B/$1.so MARLPDE_PROFILE_EVENTS=1 timeout 120 python scripts/profile_rk45.py 300 3 > $OUT/rk45_$2.log 2>&1; echo "rk45 $2: $(tail -1 $OUT/rk45_$2.log)"; }
rd() { MARLPDE_B200_LIB=$B/$1.so timeout 120 python scripts/profile_radau.py 16 0.05 > $OUT/radau_$1.log 2>&1; echo "radau $1: $(head -1 $OUT/radau_$1.log)"; }
rk lib_m0 m0
rk lib_m1 m1
rk lib_m2 m2
MARLPDE_RK45_WARP_PERM=0,1,4,3,2,5,6,7,8,9 rk lib_m1 m1_perm
MARLPDE_RK45_WARP_PERM=0,1,4,3,2,5,6,7,8,9 rk lib_m0 m0_perm
MARLPDE_RK45_BUILD=321 rk lib_m2 m2_ys
for l in lib_m0 lib_m1 lib_m2 lib_m0_mb3 lib_m0_mb5 lib_m2_mb5; do rd $l; done
for l in lib_m0 lib_m2; do
  MARLPDE_B200_LIB=$B/$l.so timeout 120 python scripts/profile_stream.py 20000 64 8 > $OUT/tiles_$l.log 2>&1; echo "tiles $l: $(tail -1 $OUT/tiles_$l.log)"
done
for l in lib_m1 lib_m2; do
  MARLPDE_B200_LIB=$B/$l.so timeout 200 python -m pytest -q -m gpu -p no:cacheprovider --timeout=120 --timeout-method=thread tests/test_gpu_rhs.py tests/test_gpu_math.py > $OUT/pytest_$l.log 2>&1
  echo "pytest $l: $(tail -1 $OUT/pytest_$l.log)"
done
echo done
