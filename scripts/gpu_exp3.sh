#!/bin/bash
set -u
OUT=gpurun_out/${1:-exp3}
mkdir -p $OUT
PT="python -m pytest -x -q -m gpu -p no:cacheprovider --timeout=150 --timeout-method=thread"
( MARLPDE_RK45_BUILD=321 timeout 400 $PT tests/test_gpu_rk45.py ) > $OUT/pytest_gpu_321.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu_321.log
timeout 200 python scripts/profile_rk45.py 1000 3 > $OUT/prof_320.log 2>&1
MARLPDE_RK45_BUILD=321 timeout 200 python scripts/profile_rk45.py 1000 3 > $OUT/prof_321.log 2>&1
MARLPDE_RK45_BUILD=321 MARLPDE_PROFILE_EVENTS=1 timeout 200 python scripts/profile_rk45.py 1000 3 > $OUT/prof_321_ev.log 2>&1
echo done
