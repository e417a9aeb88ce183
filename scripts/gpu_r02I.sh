#!/bin/bash
# r02I: where the time of a single 20 000-cell column goes on the streaming path (one launch per attempt, graph replay):
# kernel durations under ncu against the wall time per attempt of the plain run
set -u
OUT=gpurun_out/${1:-r02I}; mkdir -p $OUT
timeout 100 python scripts/profile_stream.py 20000 1 400 > $OUT/plain.log 2>&1; cat $OUT/plain.log
MARLPDE_RK45_STREAM_GRAPH=0 timeout 100 python scripts/profile_stream.py 20000 1 400 > $OUT/plain_nograph.log 2>&1; cat $OUT/plain_nograph.log
MARLPDE_RK45_STREAM_GRAPH=0 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches_b1.csv python scripts/profile_stream.py 20000 1 60 > $OUT/ncu.log 2>&1
grep -v "^==" $OUT/launches_b1.csv | awk -F'","' 'NR>1 {n[$5" "$8" "$9]++; s[$5" "$8" "$9]+=$15} END {for (k in n) printf "%s: %d launches, mean %.0f ns\n", k, n[k], s[k]/n[k]}'
echo done
