#!/bin/bash
# End-of-round run, most important outputs first: parity, smoke, bench line, reference arm, ncu launch list, ncu full
# capture of the RK45 kernel, then bench --full (sweeps to T*, implicit path, large N).
set -u
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/nvidia_smi.csv 2>&1
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout=300 --timeout-method=thread"
( time timeout 600 $PT tests ) > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
timeout 300 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err
echo "bench exit $?" >> $OUT/bench.err
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2>> $OUT/bench.err
PROF="python scripts/profile_rk45.py 300 3"
MARLPDE_PROFILE_EVENTS=1 timeout 100 $PROF > $OUT/profile_plain.log 2>&1 &&
MARLPDE_PROFILE_EVENTS=1 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches.csv $PROF > $OUT/ncu_launches.log 2>&1
# launch list of the bench command itself (the contract's "same command"): one rk45_persistent_kernel per step
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_launches_bench.log 2>&1
MARLPDE_PROFILE_EVENTS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:rk45_persistent -s 1 -c 1 -o $OUT/rk45_full $PROF > $OUT/ncu_full.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 --full --no-cpu-baseline > $OUT/bench_full.json 2>> $OUT/bench.err
echo "bench full exit $?" >> $OUT/bench.err
cat $OUT/profile_plain.log
echo done
