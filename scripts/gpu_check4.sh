#!/bin/bash
# in-tree build (RK45: lean schedule + automatic warp order; Radau: all-merged; tiles: split): full GPU parity suite,
# RK45 profile, tiles with the lean schedule for comparison
set -u
OUT=gpurun_out/${1:-check4}; mkdir -p $OUT
B=$PWD/build_ab
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout=300 --timeout-method=thread"
( time timeout 600 $PT tests ) > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -4 $OUT/pytest_gpu.log
MARLPDE_PROFILE_EVENTS=1 timeout 120 python scripts/profile_rk45.py 300 5 > $OUT/rk45_intree.log 2>&1; echo "rk45 in-tree: $(tail -2 $OUT/rk45_intree.log | tr '\n' ' ')"
MARLPDE_RK45_WARP_PERM=identity MARLPDE_PROFILE_EVENTS=1 timeout 120 python scripts/profile_rk45.py 300 5 > $OUT/rk45_intree_id.log 2>&1; echo "rk45 in-tree identity order: $(tail -2 $OUT/rk45_intree_id.log | tr '\n' ' ')"
MARLPDE_PROFILE_EVENTS=0 timeout 120 python scripts/profile_rk45.py 300 5 > $OUT/rk45_intree_noev.log 2>&1; echo "rk45 in-tree no events: $(tail -1 $OUT/rk45_intree_noev.log)"
timeout 120 python scripts/profile_stream.py 20000 64 32 > $OUT/tiles_intree.log 2>&1; echo "tiles in-tree: $(tail -1 $OUT/tiles_intree.log)"
MARLPDE_B200_LIB=$B/lib_m4.so timeout 120 python scripts/profile_stream.py 20000 64 32 > $OUT/tiles_m4.log 2>&1; echo "tiles lean: $(tail -1 $OUT/tiles_m4.log)"
timeout 120 python scripts/profile_radau.py 16 0.05 > $OUT/radau_intree.log 2>&1; echo "radau in-tree: $(head -1 $OUT/radau_intree.log)"
echo done
