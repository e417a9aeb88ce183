#!/bin/bash
# One gpurun call: GPU parity tests, bench line, ncu launch list and one full capture of the RK45 kernel.
# usage: gpurun --timeout 1800 -- bash scripts/gpu_round.sh [tag]
set -u
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/nvidia_smi.csv 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --full > $OUT/bench.json 2> $OUT/bench.err
echo "bench exit $?" >> $OUT/bench.err
timeout 300 python bench.py --steps 5 --warmup 3 --base scenario_A --no-cpu-baseline > $OUT/bench_scenarioA.json 2>> $OUT/bench.err
PROF="python scripts/profile_rk45.py 300 3"
$PROF > $OUT/profile_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches.csv $PROF > $OUT/ncu_launches.log 2>&1
$PROF > $OUT/profile_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rk45_persistent -s 1 -c 1 -o $OUT/rk45_full $PROF > $OUT/ncu_full.log 2>&1
echo done
