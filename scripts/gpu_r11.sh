#!/bin/bash
set -u
OUT=gpurun_out/${1:-r11}
mkdir -p $OUT
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout=300 --timeout-method=thread"
( time timeout 900 $PT tests ) > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu.log
timeout 200 python scripts/profile_rk45.py 1000 3 > $OUT/prof_320.log 2>&1
MARLPDE_PROFILE_EVENTS=1 timeout 200 python scripts/profile_rk45.py 1000 3 > $OUT/prof_320_ev.log 2>&1
echo done
