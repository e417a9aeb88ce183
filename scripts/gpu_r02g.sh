#!/bin/bash
# r02g: Radau kernel with the smaller instruction footprint (template flag for the model variant, predicate-bit event
# detection, event location out of line): parity suite, timing at t = 0.05 (r02a in-tree: 1.495 s) and the tail experiment.
set -u
OUT=gpurun_out/${1:-r02g}; mkdir -p $OUT
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout=400 --timeout-method=thread"
( time timeout 1500 $PT tests ) > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -12 $OUT/pytest_gpu.log
for i in 1 2; do timeout 100 python scripts/profile_radau.py 16 0.05 > $OUT/radau_005_$i.log 2>&1; head -1 $OUT/radau_005_$i.log; done
timeout 100 python scripts/profile_radau.py 4 0.05 > $OUT/radau_005_64.log 2>&1; head -1 $OUT/radau_005_64.log
timeout 600 python scripts/exp_radau_tail.py 16 > $OUT/radau_tail.log 2>&1; cat $OUT/radau_tail.log
echo done
