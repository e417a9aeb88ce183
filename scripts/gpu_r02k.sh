#!/bin/bash
# r02k: eight interleaved copies of the log / exp tables (conflict-free look-ups) in the RK45 and tile kernels
set -u
OUT=gpurun_out/${1:-r02k}; mkdir -p $OUT
for i in 1 2; do
MARLPDE_PROFILE_EVENTS=1 timeout 150 python scripts/profile_rk45.py 300 5 > $OUT/rk45_300_$i.log 2>&1; echo "rk45 300 #$i: $(tail -2 $OUT/rk45_300_$i.log | tr '\n' ' ')"
done
MARLPDE_PROFILE_EVENTS=1 timeout 150 python scripts/profile_rk45.py 3000 3 > $OUT/rk45_3000.log 2>&1; echo "rk45 3000: $(tail -2 $OUT/rk45_3000.log | tr '\n' ' ')"
timeout 120 python scripts/profile_stream.py 20000 64 96 > $OUT/stream_tmp.log 2>&1; echo "stream 20000 64: $(tail -1 $OUT/stream_tmp.log)"
timeout 120 python scripts/profile_stream.py 20000 1 256 > $OUT/stream_tmp.log 2>&1; echo "stream 20000 1: $(tail -1 $OUT/stream_tmp.log)"
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout=400 --timeout-method=thread"
( time timeout 900 $PT tests/test_gpu_rk45.py tests/test_gpu_math.py tests/test_gpu_rhs.py ) > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu.log; tail -4 $OUT/pytest_gpu.log
MARLPDE_PROFILE_EVENTS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:rk45_persistent -s 1 -c 1 -o $OUT/rk45_full python scripts/profile_rk45.py 300 3 > $OUT/ncu_rk45.log 2>&1; echo "ncu rk45 rc $?"
echo done
