#!/bin/bash
# r02t: why does the BDF kernel not finish 87 lattice columns that SciPy BDF finishes?  FD-Jacobian build, the unfinished
# columns on their own, single columns against the emulator / SciPy counts
set -u
OUT=gpurun_out/${1:-r02t}; mkdir -p $OUT
timeout 200 python scripts/diag_bdf_columns.py 0.2 1305 1360 3932 > $OUT/single_analytic.log 2>&1; cat $OUT/single_analytic.log
MARLPDE_B200_LIB=$PWD/build_ab/lib_jacfd.so timeout 200 python scripts/diag_bdf_columns.py 0.2 1305 1360 3932 > $OUT/single_fd.log 2>&1; cat $OUT/single_fd.log
timeout 300 python scripts/diag_bdf_columns.py 1.0 @build_ab/bdf_status_r02s.npz > $OUT/unfinished_analytic.log 2>&1; head -12 $OUT/unfinished_analytic.log
MARLPDE_B200_LIB=$PWD/build_ab/lib_jacfd.so timeout 300 python scripts/diag_bdf_columns.py 1.0 @build_ab/bdf_status_r02s.npz > $OUT/unfinished_fd.log 2>&1; head -12 $OUT/unfinished_fd.log
MARLPDE_B200_LIB=$PWD/build_ab/lib_jacfd.so timeout 300 python scripts/dump_implicit_status.py bdf $OUT/bdf_fd_status.npz 2>&1 | tail -1 | cut -c1-400
echo done
