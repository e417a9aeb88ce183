#!/bin/bash
set -u
OUT=gpurun_out/${1:-r10}
mkdir -p $OUT
PT="python -m pytest -x -q -m gpu -p no:cacheprovider --timeout=200 --timeout-method=thread"
( MARLPDE_RK45_BUILD=416 timeout 600 $PT tests/test_gpu_rk45.py -k "not streaming" ) > $OUT/pytest_416.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_416.log
MARLPDE_RK45_BUILD=416 timeout 200 python scripts/profile_rk45.py 1000 3 > $OUT/prof_416.log 2>&1
MARLPDE_RK45_BUILD=416 MARLPDE_PROFILE_EVENTS=1 timeout 200 python scripts/profile_rk45.py 1000 3 > $OUT/prof_416_ev.log 2>&1
timeout 200 python scripts/profile_rk45.py 1000 3 > $OUT/prof_320.log 2>&1
echo done
