#!/bin/bash
set -u
OUT=gpurun_out/${1:-r10}
mkdir -p $OUT
PT="python -m pytest -x -q -m gpu -p no:cacheprovider --timeout=200 --timeout-method=thread"
( timeout 600 $PT tests/test_gpu_rk45.py tests/test_gpu_rhs.py -k "not streaming" ) > $OUT/pytest_rk45.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_rk45.log
timeout 200 python scripts/profile_rk45.py 1000 3 > $OUT/prof_320.log 2>&1
MARLPDE_PROFILE_EVENTS=1 timeout 200 python scripts/profile_rk45.py 1000 3 > $OUT/prof_320_ev.log 2>&1
echo done
