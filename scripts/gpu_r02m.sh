#!/bin/bash
# r02m: Radau RHS inputs staged one pass ahead by cp.async (in-tree) against the same build without it (build_ab/lib_nostage.so)
set -u
OUT=gpurun_out/${1:-r02m}; mkdir -p $OUT
for i in 1 2; do
  timeout 100 python scripts/profile_radau.py 16 0.05 > $OUT/stage_$i.log 2>&1; echo "stage   : $(head -1 $OUT/stage_$i.log)"
  MARLPDE_B200_LIB=$PWD/build_ab/lib_nostage.so timeout 100 python scripts/profile_radau.py 16 0.05 > $OUT/nostage_$i.log 2>&1; echo "no stage: $(head -1 $OUT/nostage_$i.log)"
done
timeout 100 python scripts/profile_radau.py 4 0.05 > $OUT/stage_64.log 2>&1; echo "stage 64   : $(head -1 $OUT/stage_64.log)"
MARLPDE_B200_LIB=$PWD/build_ab/lib_nostage.so timeout 100 python scripts/profile_radau.py 4 0.05 > $OUT/nostage_64.log 2>&1; echo "no stage 64: $(head -1 $OUT/nostage_64.log)"
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout=400 --timeout-method=thread"
( time timeout 900 $PT tests/test_gpu_radau.py tests/test_gpu_dropin.py tests/test_gpu_lattice.py tests/test_gpu_reference_suite.py ) > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu.log; tail -4 $OUT/pytest_gpu.log
timeout 200 python scripts/profile_radau.py 16 1.0 > $OUT/stage_tstar.log 2>&1; echo "stage T*: $(head -1 $OUT/stage_tstar.log)"
echo done
