#!/bin/bash
# r02z: PAIRED shape of the on-chip RK45 kernel (cluster of two CTAs sharing a 7th column), first GPU run after the
# emulator: bit-for-bit against the classic shape on 4096 / 8192 / 64 / 432 / 448 columns with timings, then the RK45
# parity tests with the paired shape forced
set -u
OUT=gpurun_out/${1:-r02z}; mkdir -p $OUT
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout=400 --timeout-method=thread"
timeout 420 python scripts/exp_rk45_pair.py 3000 3 > $OUT/pair_vs_solo.log 2>&1; echo "exit $?" >> $OUT/pair_vs_solo.log; cat $OUT/pair_vs_solo.log
( time MARLPDE_RK45_SHAPE=pair timeout 600 $PT tests/test_gpu_rk45.py tests/test_gpu_lattice.py -k "not radau and not bdf" ) > $OUT/pytest_pair.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_pair.log; tail -8 $OUT/pytest_pair.log
echo done
