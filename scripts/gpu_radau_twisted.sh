#!/bin/bash
# two-ended block-Thomas (in-tree build) vs the one-directional sweeps (build_ab/lib_head.so): parity, then timing
set -u
OUT=gpurun_out/${1:-radau_tw}; mkdir -p $OUT
timeout 400 python -m pytest -q -m gpu -p no:cacheprovider --timeout=300 --timeout-method=thread tests/test_gpu_radau.py tests/test_gpu_dropin.py > $OUT/pytest.log 2>&1
echo "pytest exit $?" >> $OUT/pytest.log; tail -3 $OUT/pytest.log
timeout 100 python scripts/profile_radau.py 16 0.05 > $OUT/new.log 2>&1; echo "new 4096: $(head -1 $OUT/new.log)"
MARLPDE_B200_LIB=$PWD/build_ab/lib_head.so timeout 100 python scripts/profile_radau.py 16 0.05 > $OUT/head.log 2>&1; echo "head 4096: $(head -1 $OUT/head.log)"
timeout 100 python scripts/profile_radau.py 4 0.05 > $OUT/new_64.log 2>&1; echo "new 64: $(head -1 $OUT/new_64.log)"
MARLPDE_B200_LIB=$PWD/build_ab/lib_head.so timeout 100 python scripts/profile_radau.py 4 0.05 > $OUT/head_64.log 2>&1; echo "head 64: $(head -1 $OUT/head_64.log)"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
echo done
