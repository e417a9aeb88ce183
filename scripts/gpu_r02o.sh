#!/bin/bash
# r02o: implicit sweep ordered by the implicit-path cost estimate (bench block)
set -u
OUT=gpurun_out/${1:-r02o}; mkdir -p $OUT
( time timeout 900 python bench.py --steps 3 --no-large-n --no-cpu-baseline --no-equal-load --skip-rk45-tstar ) > $OUT/bench.json 2> $OUT/bench.err
python - <<'PY' $OUT/bench.json
import json,sys
j=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print('value',j['value'])
for b,v in j.get('implicit_time_to_Tstar',{}).items(): print(b,v['seconds'],v['finished'],v.get('repeat_sweep_longest_first'))
PY
echo done
