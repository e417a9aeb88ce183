#!/bin/bash
# experiment: two-cells-per-thread RK45 kernel, both builds; every step has its own short timeout
set -u
OUT=gpurun_out/${1:-exp2}
mkdir -p $OUT
PT="python -m pytest -x -q -m gpu -p no:cacheprovider --timeout=150 --timeout-method=thread"
( time timeout 600 $PT tests ) > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu.log
timeout 200 python scripts/profile_rk45.py 1000 3 > $OUT/prof_320.log 2>&1
MARLPDE_RK45_BUILD=400 timeout 200 python scripts/profile_rk45.py 1000 3 > $OUT/prof_400.log 2>&1
MARLPDE_PROFILE_EVENTS=1 timeout 200 python scripts/profile_rk45.py 1000 3 > $OUT/prof_320_ev.log 2>&1
timeout 200 python scripts/profile_rk45.py 1000 3 scenario_A > $OUT/prof_320_A.log 2>&1
MARLPDE_RK45_BUILD=400 timeout 200 python scripts/profile_rk45.py 1000 3 scenario_A > $OUT/prof_400_A.log 2>&1
( MARLPDE_RK45_BUILD=400 timeout 400 $PT tests/test_gpu_rk45.py ) > $OUT/pytest_gpu_400.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu_400.log
echo done
