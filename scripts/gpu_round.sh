#!/bin/bash
# One gpurun call: GPU parity tests, bench lines, ncu launch lists and full captures of the hot kernels.
# usage: gpurun --timeout 2700 -- bash scripts/gpu_round.sh [tag]
set -u
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/nvidia_smi.csv 2>&1
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout=300 --timeout-method=thread"
( time timeout 900 $PT tests ) > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1
timeout 1200 python bench.py --steps 5 --warmup 3 --full > $OUT/bench.json 2> $OUT/bench.err
echo "bench exit $?" >> $OUT/bench.err
timeout 300 python bench.py --steps 5 --warmup 3 --base scenario_A --no-cpu-baseline > $OUT/bench_scenarioA.json 2>> $OUT/bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2>> $OUT/bench.err
PROF="python scripts/profile_rk45.py 300 3"
MARLPDE_PROFILE_EVENTS=1 timeout 200 $PROF > $OUT/profile_plain.log 2>&1 &&
MARLPDE_PROFILE_EVENTS=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches.csv $PROF > $OUT/ncu_launches.log 2>&1
MARLPDE_PROFILE_EVENTS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:rk45_persistent -s 1 -c 1 -o $OUT/rk45_full $PROF > $OUT/ncu_full.log 2>&1
SPROF="python scripts/profile_stream.py 20000 64 8"
timeout 200 $SPROF > $OUT/tiles_plain.log 2>&1 &&
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 80 --csv --log-file $OUT/tiles_launches.csv $SPROF > $OUT/ncu_tiles_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_attempt -s 5 -c 1 -o $OUT/tile_full $SPROF > $OUT/ncu_tile_full.log 2>&1
RPROF="python scripts/profile_radau.py 16 0.01"
timeout 200 $RPROF > $OUT/radau_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:radau_kernel -s 1 -c 1 -o $OUT/radau_full $RPROF > $OUT/ncu_radau_full.log 2>&1
echo done
