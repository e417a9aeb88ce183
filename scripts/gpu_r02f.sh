#!/bin/bash
# r02f: parity suite (streaming events, model variant, lattice columns, tile builds), streaming timings per window size,
# ncu --set full captures of the three hot kernels (each after its plain run exited 0).
set -u
OUT=gpurun_out/${1:-r02f}; mkdir -p $OUT
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout=400 --timeout-method=thread"
( time timeout 1500 $PT tests ) > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -25 $OUT/pytest_gpu.log
for t in small large; do for nb in "20000 1" "5000 1" "2000 8" "20000 8"; do
  MARLPDE_RK45_TILE=$t timeout 120 python scripts/profile_stream.py $nb 256 > $OUT/stream_tmp.log 2>&1; echo "stream $nb tile=$t: $(tail -1 $OUT/stream_tmp.log)"
done; done
MARLPDE_RK45_TILE_TMA=1 timeout 120 python scripts/profile_stream.py 20000 64 96 > $OUT/stream_tmp.log 2>&1; echo "stream 20000 64 tma=1: $(tail -1 $OUT/stream_tmp.log)"
timeout 120 python scripts/profile_stream.py 20000 64 96 > $OUT/stream_tmp.log 2>&1; echo "stream 20000 64 tma=0: $(tail -1 $OUT/stream_tmp.log)"
PROF="python scripts/profile_rk45.py 300 3"
MARLPDE_PROFILE_EVENTS=1 timeout 100 $PROF > $OUT/profile_plain.log 2>&1 &&
MARLPDE_PROFILE_EVENTS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:rk45_persistent -s 1 -c 1 -o $OUT/rk45_full $PROF > $OUT/ncu_rk45.log 2>&1; echo "ncu rk45 rc $?"
timeout 100 python scripts/profile_radau.py 16 0.01 > $OUT/radau_plain.log 2>&1 && tail -2 $OUT/radau_plain.log &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:radau_kernel -s 1 -c 1 -o $OUT/radau_full python scripts/profile_radau.py 16 0.01 > $OUT/ncu_radau.log 2>&1; echo "ncu radau rc $?"
timeout 100 python scripts/profile_stream.py 20000 64 24 > $OUT/stream_plain.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tile_attempt -s 30 -c 1 -o $OUT/tile_full python scripts/profile_stream.py 20000 64 24 > $OUT/ncu_tile.log 2>&1; echo "ncu tile rc $?"
echo done
