#!/bin/bash
set -u
OUT=gpurun_out/${1:-sanitize}
mkdir -p $OUT
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 120 python scripts/sanitize_small.py all > $OUT/plain.log 2>&1
for w in rhs rk45 tiles stages radau; do
  timeout 600 $CS --tool memcheck --error-exitcode 9 python scripts/sanitize_small.py $w > $OUT/memcheck_$w.log 2>&1; echo "exit $?" >> $OUT/memcheck_$w.log
done
for w in rk45 tiles radau; do
  timeout 900 $CS --tool racecheck --racecheck-report all --error-exitcode 9 python scripts/sanitize_small.py $w > $OUT/racecheck_$w.log 2>&1; echo "exit $?" >> $OUT/racecheck_$w.log
done
echo done
