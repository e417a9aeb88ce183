#!/bin/bash
# r02z2: where the paired shape loses its time — the same comparison with the relaxed-synchronisation build
# (build_ab/lib_pair_relaxed.so, -DMARLPDE_PAIR_SYNC=1: no MEMBAR.GPU on arrive, no CCTL.IVALL on the stage waits)
set -u
OUT=gpurun_out/${1:-r02z2}; mkdir -p $OUT
export MARLPDE_EXP_LATTICES="16,16,16;4,4,4"
timeout 300 python scripts/exp_rk45_pair.py 3000 2 > $OUT/pair_strict.log 2>&1; echo "exit $?" >> $OUT/pair_strict.log; cat $OUT/pair_strict.log
MARLPDE_B200_LIB=$PWD/build_ab/lib_pair_relaxed.so timeout 300 python scripts/exp_rk45_pair.py 3000 2 > $OUT/pair_relaxed.log 2>&1; echo "exit $?" >> $OUT/pair_relaxed.log; cat $OUT/pair_relaxed.log
echo done
