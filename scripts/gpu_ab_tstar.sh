#!/bin/bash
# A/B of two library builds on the same box: phase-resolved RK45 sweep to T*.
OUT=gpurun_out/ab_tstar; mkdir -p $OUT
MARLPDE_B200_LIB=$PWD/build_ab/libold.so timeout 400 python scripts/diag_tstar.py 888 50000 > $OUT/old.log 2>&1
timeout 400 python scripts/diag_tstar.py 888 50000 > $OUT/new.log 2>&1
tail -1 $OUT/old.log $OUT/new.log
