#!/bin/bash
set -u
OUT=gpurun_out/${1:-r3}
mkdir -p $OUT
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout=300 --timeout-method=thread"
( time timeout 900 $PT tests/test_gpu_rk45.py -k "streaming or unsupported" ) > $OUT/pytest_stream.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_stream.log
( time timeout 900 $PT tests/test_gpu_radau.py ) > $OUT/pytest_radau.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_radau.log
timeout 600 python scripts/diag_radau.py > $OUT/diag.log 2>&1
echo done
