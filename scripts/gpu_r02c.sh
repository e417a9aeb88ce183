#!/bin/bash
# r02c: full GPU parity suite, smoke, the default bench line (time-to-T*, implicit sweep, large N, equal load) and the reference arm.
set -u
OUT=gpurun_out/${1:-r02c}; mkdir -p $OUT
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout=300 --timeout-method=thread"
( time timeout 900 $PT tests ) > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
( time timeout 900 python bench.py ) > $OUT/bench.json 2> $OUT/bench.err
echo "bench exit $?" >> $OUT/bench.err
tail -5 $OUT/bench.err
( time timeout 300 python bench.py --impl reference --steps 2 --warmup 1 ) > $OUT/bench_reference.json 2>> $OUT/bench.err
python - <<'PY' $OUT/bench.json
import json,sys
try:
    j=json.loads(open(sys.argv[1]).readline())
    print({k:(v if not isinstance(v,dict) else '...') for k,v in j.items()})
    print('equal_load',j.get('equal_load',{}).get('value'))
    print('tstar',{k:v for k,v in j.get('time_to_Tstar',{}).items() if k in('seconds','finished','status_histogram','step_attempts')})
    for b,v in j.get('implicit_time_to_Tstar',{}).items(): print(b,v['seconds'],v['finished'],v['roofline']['frac'],v.get('cpu_baseline'))
    for b,v in j.get('large_n_streaming',{}).items(): print(b,v['column_steps_per_s'],v['roofline']['frac'],v['roofline_hbm']['frac'])
    print(j['roofline']); print(j['cpu_baseline']); print(j['e2e'])
except Exception as e: print('parse failed',e)
PY
echo done
