#!/bin/bash
set -u
OUT=gpurun_out/${1:-r5}
mkdir -p $OUT
timeout 300 python scripts/profile_radau.py 16 0.05 > $OUT/radau_4096.log 2>&1
timeout 300 python scripts/profile_radau.py 4 0.05 > $OUT/radau_64.log 2>&1
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout=300 --timeout-method=thread"
( time timeout 900 $PT tests/test_gpu_radau.py tests/test_gpu_dropin.py ) > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu.log
echo done
