#!/bin/bash
# r02n: Radau outward sweep with precomputed X records (in-tree) against the same build without them (build_ab/lib_noxrec.so)
set -u
OUT=gpurun_out/${1:-r02n}; mkdir -p $OUT
for i in 1 2; do
  timeout 100 python scripts/profile_radau.py 16 0.05 > $OUT/xrec_$i.log 2>&1; echo "xrec   : $(head -1 $OUT/xrec_$i.log)"
  MARLPDE_B200_LIB=$PWD/build_ab/lib_noxrec.so timeout 100 python scripts/profile_radau.py 16 0.05 > $OUT/noxrec_$i.log 2>&1; echo "no xrec: $(head -1 $OUT/noxrec_$i.log)"
done
timeout 100 python scripts/profile_radau.py 4 0.05 > $OUT/xrec_64.log 2>&1; echo "xrec 64   : $(head -1 $OUT/xrec_64.log)"
MARLPDE_B200_LIB=$PWD/build_ab/lib_noxrec.so timeout 100 python scripts/profile_radau.py 4 0.05 > $OUT/noxrec_64.log 2>&1; echo "no xrec 64: $(head -1 $OUT/noxrec_64.log)"
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout=400 --timeout-method=thread"
( time timeout 900 $PT tests/test_gpu_radau.py tests/test_gpu_dropin.py tests/test_gpu_lattice.py tests/test_gpu_reference_suite.py ) > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu.log; tail -4 $OUT/pytest_gpu.log
timeout 200 python scripts/profile_radau.py 16 1.0 > $OUT/xrec_tstar.log 2>&1; echo "xrec T*: $(head -1 $OUT/xrec_tstar.log)"
echo done
