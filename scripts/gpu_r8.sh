#!/bin/bash
set -u
OUT=gpurun_out/${1:-r8}
mkdir -p $OUT
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout=300 --timeout-method=thread"
( time timeout 900 $PT tests/test_gpu_rk45.py -k streaming ) > $OUT/pytest_stream.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_stream.log
for a in "20000 64 32" "2000 64 64" "20000 1 64" "5000 8 64"; do timeout 200 python scripts/profile_stream.py $a >> $OUT/stream_tiles.log 2>&1; done
echo done
