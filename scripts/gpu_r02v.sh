#!/bin/bash
# r02v: "analytic first, one finite-difference retry" Jacobian against the default (finite-difference diagonal blocks), both
# implicit kernels: t = 0.05 timing and the sweep to T* (status dumps); BDF kernel at 2 / 3 / 4 resident CTAs per SM
set -u
OUT=gpurun_out/${1:-r02v}; mkdir -p $OUT
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout=400 --timeout-method=thread"
( time timeout 900 $PT tests/test_gpu_bdf.py tests/test_gpu_radau.py ) > $OUT/pytest_implicit.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_implicit.log; tail -5 $OUT/pytest_implicit.log
for m in radau bdf; do for j in analytic fd; do
  timeout 300 python scripts/profile_implicit.py $m 16 0.05 0 $j > $OUT/${m}_${j}_4096_t005.log 2>&1; echo "$j $(head -1 $OUT/${m}_${j}_4096_t005.log)"
done; done
for m in radau bdf; do
  timeout 300 python scripts/dump_implicit_status.py $m $OUT/${m}_analytic_status.npz analytic 2>&1 | tail -1 | cut -c1-300
done
for mb in 2 4; do
  MARLPDE_B200_LIB=$PWD/build_ab/lib_bdf_mb$mb.so timeout 300 python scripts/profile_implicit.py bdf 16 0.05 > $OUT/bdf_mb${mb}_4096_t005.log 2>&1; echo "minblocks $mb: $(head -1 $OUT/bdf_mb${mb}_4096_t005.log)"
  MARLPDE_B200_LIB=$PWD/build_ab/lib_bdf_mb$mb.so timeout 300 python scripts/profile_implicit.py bdf 4 0.05 > $OUT/bdf_mb${mb}_64_t005.log 2>&1; echo "minblocks $mb: $(head -1 $OUT/bdf_mb${mb}_64_t005.log)"
done
echo done
