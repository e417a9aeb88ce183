#!/bin/bash
# r02s: implicit kernels after the analytic Jacobian / compact BDF records / fused BDF residual:
#   parity (BDF, Radau incl. the device Jacobian probe, lattice, drop-in), A/B analytic vs finite-difference Jacobian
#   (build_ab/lib_jacfd.so), per-column status dumps of both implicit sweeps to T*, ncu --set full of the BDF kernel
set -u
OUT=gpurun_out/${1:-r02s}; mkdir -p $OUT
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout=400 --timeout-method=thread"
( time timeout 900 $PT tests/test_gpu_bdf.py tests/test_gpu_radau.py tests/test_gpu_lattice.py tests/test_gpu_dropin.py tests/test_gpu_reference_suite.py ) > $OUT/pytest_implicit.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_implicit.log; tail -12 $OUT/pytest_implicit.log
for rep in 1 2; do
  for m in radau bdf; do
    timeout 300 python scripts/profile_implicit.py $m 16 0.05 > $OUT/${m}_4096_t005_$rep.log 2>&1; echo "analytic $m 4096 t=0.05: $(head -1 $OUT/${m}_4096_t005_$rep.log)"
    MARLPDE_B200_LIB=$PWD/build_ab/lib_jacfd.so timeout 300 python scripts/profile_implicit.py $m 16 0.05 > $OUT/${m}_fd_4096_t005_$rep.log 2>&1; echo "FD       $m 4096 t=0.05: $(head -1 $OUT/${m}_fd_4096_t005_$rep.log)"
  done
done
for m in radau bdf; do
  timeout 300 python scripts/profile_implicit.py $m 4 0.05 > $OUT/${m}_64_t005.log 2>&1; echo "analytic $m 64: $(head -1 $OUT/${m}_64_t005.log)"
  MARLPDE_B200_LIB=$PWD/build_ab/lib_jacfd.so timeout 300 python scripts/profile_implicit.py $m 4 0.05 > $OUT/${m}_fd_64_t005.log 2>&1; echo "FD       $m 64: $(head -1 $OUT/${m}_fd_64_t005.log)"
done
timeout 300 python scripts/dump_implicit_status.py radau $OUT/radau_status.npz 2>&1 | tail -1
MARLPDE_B200_LIB=$PWD/build_ab/lib_jacfd.so timeout 300 python scripts/dump_implicit_status.py radau $OUT/radau_fd_status.npz 2>&1 | tail -1
timeout 300 python scripts/dump_implicit_status.py bdf $OUT/bdf_status.npz 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bdf_kernel -c 1 -o $OUT/bdf_full python scripts/profile_implicit.py bdf 16 0.01 > $OUT/ncu_bdf.log 2>&1; tail -2 $OUT/ncu_bdf.log
echo done
