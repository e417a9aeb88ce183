#!/bin/bash
# r02q: the tile kernel that closes the previous attempt in its prologue (one launch per attempt): streaming parity tests and
# the large-N timings of the bench (20 000 x 64, 2 000 x 64, 20 000 x 1) + launch list of a short streaming run
set -u
OUT=gpurun_out/${1:-r02q}; mkdir -p $OUT
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout=400 --timeout-method=thread"
( time timeout 600 $PT tests/test_gpu_rk45.py ) > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu.log; tail -5 $OUT/pytest_gpu.log
for cfg in "20000 64 48" "2000 64 200" "20000 1 400" "20000 8 100" "5000 8 200"; do
  timeout 120 python scripts/profile_stream.py $cfg > $OUT/stream_${cfg// /_}.log 2>&1; echo "$cfg: $(tail -1 $OUT/stream_${cfg// /_}.log)"
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/launches_stream.csv python scripts/profile_stream.py 20000 64 12 > $OUT/ncu_stream.log 2>&1
echo done
