#!/bin/bash
set -u
OUT=gpurun_out/${1:-ncu_radau3}
mkdir -p $OUT
timeout 300 python scripts/profile_radau.py 16 0.01 > $OUT/plain_4096.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:radau_kernel -s 1 -c 1 -o $OUT/radau_full_4096 python scripts/profile_radau.py 16 0.01 > $OUT/ncu_full.log 2>&1
echo done
