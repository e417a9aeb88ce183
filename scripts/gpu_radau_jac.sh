#!/bin/bash
# hybrid Jacobian (analytic off-diagonal blocks + 5 FD passes): parity tests, then timing vs the previous build
set -u
OUT=gpurun_out/${1:-radau_jac}; mkdir -p $OUT
timeout 600 python -m pytest -q -m gpu -p no:cacheprovider --timeout=300 --timeout-method=thread tests/test_gpu_radau.py tests/test_gpu_dropin.py > $OUT/pytest.log 2>&1
echo "pytest exit $?" >> $OUT/pytest.log
timeout 120 python scripts/profile_radau.py 16 0.05 > $OUT/new.log 2>&1
MARLPDE_B200_LIB=$PWD/build_ab/lib_fact.so timeout 120 python scripts/profile_radau.py 16 0.05 > $OUT/fact.log 2>&1
timeout 120 python scripts/profile_radau.py 4 0.05 > $OUT/new_64.log 2>&1
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1
grep -h "columns to" $OUT/*.log; tail -3 $OUT/pytest.log; tail -1 $OUT/smoke.log
