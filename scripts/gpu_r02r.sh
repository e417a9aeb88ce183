#!/bin/bash
# r02r: first GPU contact of the BDF kernel (csrc/bdf_batch.cu) + the implicit kernels after the split into
# implicit_common.cuh: BDF and Radau parity tests, timings of both on the benchmark lattice
set -u
OUT=gpurun_out/${1:-r02r}; mkdir -p $OUT
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout=400 --timeout-method=thread"
( time timeout 900 $PT tests/test_gpu_bdf.py ) > $OUT/pytest_bdf.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_bdf.log; tail -15 $OUT/pytest_bdf.log
( time timeout 900 $PT tests/test_gpu_radau.py tests/test_gpu_lattice.py tests/test_gpu_dropin.py ) > $OUT/pytest_radau.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_radau.log; tail -5 $OUT/pytest_radau.log
for m in radau bdf; do
  timeout 300 python scripts/profile_implicit.py $m 16 0.05 > $OUT/${m}_4096_t005.log 2>&1; echo "$m 4096 t=0.05: $(head -1 $OUT/${m}_4096_t005.log)"
  timeout 300 python scripts/profile_implicit.py $m 4 0.05 > $OUT/${m}_64_t005.log 2>&1; echo "$m 64 t=0.05: $(head -1 $OUT/${m}_64_t005.log)"
done
timeout 600 python scripts/profile_implicit.py bdf 16 1.0 > $OUT/bdf_4096_tstar.log 2>&1; echo "bdf 4096 T*: $(head -1 $OUT/bdf_4096_tstar.log)"
echo done
