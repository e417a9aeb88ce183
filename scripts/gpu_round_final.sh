#!/bin/bash
# Full round, most important outputs first: parity suite, smoke, default bench line (all blocks), reference arm, ncu launch
# lists of the profile command and of the bench command itself.
set -u
TAG=${1:-r02}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/nvidia_smi.csv 2>&1
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout=400 --timeout-method=thread"
( time timeout 1500 $PT tests ) > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -6 $OUT/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
( time timeout 900 python bench.py ) > $OUT/bench.json 2> $OUT/bench.err
echo "bench exit $?" >> $OUT/bench.err
tail -4 $OUT/bench.err
( time timeout 300 python bench.py --impl reference --steps 2 --warmup 1 ) > $OUT/bench_reference.json 2>> $OUT/bench.err
PROF="python scripts/profile_rk45.py 300 3"
MARLPDE_PROFILE_EVENTS=1 timeout 100 $PROF > $OUT/profile_plain.log 2>&1 &&
MARLPDE_PROFILE_EVENTS=1 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches.csv $PROF > $OUT/ncu_launches.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-tstar --no-large-n --no-equal-load > $OUT/ncu_launches_bench.log 2>&1
python - <<'PY' $OUT/bench.json
import json,sys
try:
    j=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    print({k:(v if not isinstance(v,dict) else '...') for k,v in j.items()})
    print('equal_load',j.get('equal_load',{}).get('value'))
    print('tstar',{k:v for k,v in j.get('time_to_Tstar',{}).items() if k in('seconds','finished','status_histogram','step_attempts','column_steps_per_s')})
    for b,v in j.get('implicit_time_to_Tstar',{}).items(): print(b,v['seconds'],v['finished'],v['roofline']['frac'],v.get('repeat_sweep_longest_first'),{k:x for k,x in (v.get('cpu_baseline') or {}).items() if k.startswith('seconds')})
    for b,v in j.get('bdf_time_to_Tstar',{}).items(): print('bdf',b,v['seconds'],v['finished'],v['roofline']['frac'],{k:x for k,x in (v.get('cpu_baseline') or {}).items() if k.startswith('seconds')})
    for b,v in j.get('large_n_streaming',{}).items(): print(b,v['column_steps_per_s'],v['roofline']['frac'],v['roofline_hbm']['frac'])
    print(j['roofline']); print(j['cpu_baseline']); print(j['e2e'])
except Exception as e: print('parse failed',e)
PY
echo done
