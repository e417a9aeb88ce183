#!/bin/bash
# r02G: executed warp instructions of the on-chip RK45 kernel, packed-position build against the previous one
set -u
OUT=gpurun_out/${1:-r02G}; mkdir -p $OUT
export MARLPDE_PROFILE_EVENTS=1
M=smsp__inst_executed.sum,gpu__time_duration.sum,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active
timeout 200 ncu --metrics $M --clock-control none -k regex:rk45_persistent -s 1 -c 1 --csv --log-file $OUT/new.csv python scripts/profile_rk45.py 300 3 > $OUT/new.log 2>&1
MARLPDE_B200_LIB=$PWD/build_ab/lib_head.so timeout 200 ncu --metrics $M --clock-control none -k regex:rk45_persistent -s 1 -c 1 --csv --log-file $OUT/head.csv python scripts/profile_rk45.py 300 3 > $OUT/head.log 2>&1
for v in new head; do echo "== $v"; grep -v "^==" $OUT/$v.csv | cut -d, -f13- ; done
echo done
