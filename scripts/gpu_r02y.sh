#!/bin/bash
# r02y: TEAM columns of the Radau kernel (two warps per column for the first K columns of a cost-ordered sweep):
# parity, single-column / small-batch latency, the 4096-column sweep to T* for K = 0 / 128 / 256 / 400 / 600
set -u
OUT=gpurun_out/${1:-r02y}; mkdir -p $OUT
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout=400 --timeout-method=thread"
( time timeout 900 $PT tests/test_gpu_radau.py tests/test_gpu_dropin.py tests/test_gpu_lattice.py ) > $OUT/pytest_implicit.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_implicit.log; tail -5 $OUT/pytest_implicit.log
timeout 600 python scripts/exp_radau_team.py > $OUT/team.log 2>&1; cat $OUT/team.log
echo done
