#!/bin/bash
# r02E: ncu --set full (with source) of the on-chip RK45 kernel on the final tree: 4096 columns x 300 attempts, events on
set -u
OUT=gpurun_out/${1:-r02E}; mkdir -p $OUT
PROF="python scripts/profile_rk45.py 300 3"
MARLPDE_PROFILE_EVENTS=1 timeout 100 $PROF > $OUT/profile_plain.log 2>&1 && cat $OUT/profile_plain.log &&
MARLPDE_PROFILE_EVENTS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:rk45_persistent -s 1 -c 1 -o $OUT/rk45_full $PROF > $OUT/ncu_rk45.log 2>&1; echo "ncu rk45 rc $?"
ls -la $OUT
echo done
