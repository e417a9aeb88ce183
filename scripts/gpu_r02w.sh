#!/bin/bash
# r02w: analytic Jacobian with the porosity column of cells ON a switching surface taken from num_jac's difference quotient
# (kSwitchTol) against the default finite-difference diagonal blocks: parity, t = 0.05 timing, sweeps to T* (status dumps)
set -u
OUT=gpurun_out/${1:-r02w}; mkdir -p $OUT
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout=400 --timeout-method=thread"
( time timeout 900 $PT tests/test_gpu_bdf.py tests/test_gpu_radau.py ) > $OUT/pytest_implicit.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_implicit.log; tail -5 $OUT/pytest_implicit.log
for m in radau bdf; do for j in analytic fd; do
  timeout 300 python scripts/profile_implicit.py $m 16 0.05 0 $j > $OUT/${m}_${j}_4096_t005.log 2>&1; echo "$j $(head -1 $OUT/${m}_${j}_4096_t005.log)"
done; done
for m in radau bdf; do
  timeout 300 python scripts/dump_implicit_status.py $m $OUT/${m}_analytic_status.npz analytic 2>&1 | tail -1 | cut -c1-300
done
echo done
