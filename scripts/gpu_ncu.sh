#!/bin/bash
# ncu of the persistent RK45 kernel: launch list + one full capture (after a plain run exited 0)
set -u
OUT=gpurun_out/${1:-ncu}
mkdir -p $OUT
PROF="python scripts/profile_rk45.py 300 3"
timeout 200 $PROF > $OUT/profile_plain.log 2>&1 || exit 1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches.csv $PROF > $OUT/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rk45_persistent -s 1 -c 1 -o $OUT/rk45_full $PROF > $OUT/ncu_full.log 2>&1
echo done
