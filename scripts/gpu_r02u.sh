#!/bin/bash
# r02u: implicit kernels with the Jacobian as a run-time choice (default: finite-difference diagonal blocks) and the BDF
# predictor fix: parity suite, both sweeps to T* with both Jacobians (status dumps), ncu --set full of the BDF kernel
set -u
OUT=gpurun_out/${1:-r02u}; mkdir -p $OUT
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout=400 --timeout-method=thread"
( time timeout 900 $PT tests/test_gpu_bdf.py tests/test_gpu_radau.py tests/test_gpu_lattice.py tests/test_gpu_dropin.py tests/test_gpu_reference_suite.py ) > $OUT/pytest_implicit.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_implicit.log; tail -8 $OUT/pytest_implicit.log
for m in radau bdf; do for j in fd analytic; do
  timeout 300 python scripts/profile_implicit.py $m 16 0.05 0 $j > $OUT/${m}_${j}_4096_t005.log 2>&1; echo "$(head -1 $OUT/${m}_${j}_4096_t005.log)"
  timeout 300 python scripts/dump_implicit_status.py $m $OUT/${m}_${j}_status.npz $j 2>&1 | tail -1 | cut -c1-300
done; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bdf_kernel --launch-skip 1 -c 1 -o $OUT/bdf_full python scripts/profile_implicit.py bdf 16 0.01 > $OUT/ncu_bdf.log 2>&1; tail -2 $OUT/ncu_bdf.log
echo done
