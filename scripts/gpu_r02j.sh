#!/bin/bash
# r02j: hybrid work queue on the sweep to T* (bench block only), lattice test
set -u
OUT=gpurun_out/${1:-r02j}; mkdir -p $OUT
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout=400 --timeout-method=thread"
( time timeout 900 $PT tests/test_gpu_lattice.py tests/test_gpu_rk45.py ) > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu.log; tail -4 $OUT/pytest_gpu.log
( time timeout 900 python bench.py --steps 3 --no-large-n --no-cpu-baseline --no-equal-load ) > $OUT/bench.json 2> $OUT/bench.err
python - <<'PY' $OUT/bench.json
import json,sys
j=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print('value',j['value'])
print('tstar',{k:v for k,v in j.get('time_to_Tstar',{}).items() if k in('seconds','finished','status_histogram','step_attempts','column_steps_per_s','method')})
for b,v in j.get('implicit_time_to_Tstar',{}).items(): print(b,v['seconds'],v['finished'],v.get('repeat_sweep_longest_first'))
PY
echo done
