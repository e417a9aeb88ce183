#!/bin/bash
set -u
OUT=gpurun_out/${1:-r6}
mkdir -p $OUT
MARLPDE_LPT=1 timeout 400 python scripts/profile_radau.py 16 0.05 > $OUT/radau_lpt.log 2>&1
for a in "20000 64 16" "2000 64 64" "20000 1 64" "5000 8 64"; do timeout 200 python scripts/profile_stream.py $a >> $OUT/stream.log 2>&1; done
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout=300 --timeout-method=thread"
( time timeout 900 $PT tests/test_gpu_rk45.py -k streaming ) > $OUT/pytest_stream.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_stream.log
echo done
