#!/bin/bash
set -u
OUT=gpurun_out/${1:-radau1}
mkdir -p $OUT
timeout 600 python scripts/diag_radau.py > $OUT/diag.log 2>&1
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout=240 --timeout-method=thread"
( time timeout 900 $PT tests/test_gpu_radau.py ) > $OUT/pytest_radau.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_radau.log
echo done
