#!/bin/bash
# r02d: after the clean-up / model variant / streaming events / async host entry points: parity suite, RK45 timing against
# r02b (22.6-23.3 M at 300 attempts, 25.15 M at 3000 with quanta), streaming timing with graph replay on (default) / off.
set -u
OUT=gpurun_out/${1:-r02d}; mkdir -p $OUT
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout=300 --timeout-method=thread"
( time timeout 1200 $PT tests ) > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -25 $OUT/pytest_gpu.log
for i in 1 2; do
MARLPDE_PROFILE_EVENTS=1 timeout 150 python scripts/profile_rk45.py 300 5 > $OUT/rk45_300_$i.log 2>&1; echo "rk45 300 #$i: $(tail -2 $OUT/rk45_300_$i.log | tr '\n' ' ')"
done
MARLPDE_PROFILE_EVENTS=1 timeout 150 python scripts/profile_rk45.py 3000 3 > $OUT/rk45_3000.log 2>&1; echo "rk45 3000: $(tail -2 $OUT/rk45_3000.log | tr '\n' ' ')"
for g in 1 0; do for nb in "20000 1" "2000 8" "20000 64"; do
  MARLPDE_RK45_STREAM_GRAPH=$g timeout 120 python scripts/profile_stream.py $nb 256 > $OUT/stream_tmp.log 2>&1; echo "stream $nb graph=$g: $(tail -1 $OUT/stream_tmp.log)"
done; done
timeout 300 python bench.py --steps 5 --no-tstar --no-large-n --no-cpu-baseline > $OUT/bench_short.json 2> $OUT/bench_short.err; echo "bench rc $?"; cut -c1-300 $OUT/bench_short.json
echo done
