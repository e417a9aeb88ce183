#!/bin/bash
# Lean variant of gpu_round.sh (~6 min of box time): GPU parity tests, smoke, bench lines (no explicit sweep to T*),
# ncu launch list of the bench step, ncu full capture of the Radau kernel.
# usage: gpurun --timeout 600 -- bash scripts/gpu_round_lean.sh [tag]
set -u
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/nvidia_smi.csv 2>&1
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout=300 --timeout-method=thread"
( time timeout 600 $PT tests ) > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1
timeout 400 python bench.py --steps 5 --warmup 3 --full --skip-rk45-tstar > $OUT/bench.json 2> $OUT/bench.err
echo "bench exit $?" >> $OUT/bench.err
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2>> $OUT/bench.err
PROF="python scripts/profile_rk45.py 300 3"
MARLPDE_PROFILE_EVENTS=1 timeout 100 $PROF > $OUT/profile_plain.log 2>&1 &&
MARLPDE_PROFILE_EVENTS=1 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches.csv $PROF > $OUT/ncu_launches.log 2>&1
RPROF="python scripts/profile_radau.py 16 0.01"
timeout 100 python scripts/profile_radau.py 16 0.05 > $OUT/radau_plain_005.log 2>&1
timeout 100 $RPROF > $OUT/radau_plain.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:radau_kernel -s 1 -c 1 -o $OUT/radau_full $RPROF > $OUT/ncu_radau_full.log 2>&1
tail -3 $OUT/pytest_gpu.log; tail -1 $OUT/smoke.log; cat $OUT/radau_plain_005.log
echo done
