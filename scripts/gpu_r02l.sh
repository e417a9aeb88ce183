#!/bin/bash
# r02l: ncu of the Radau kernel after the footprint work; window-size rule between half and all of the SMs
set -u
OUT=gpurun_out/${1:-r02l}; mkdir -p $OUT
timeout 100 python scripts/profile_radau.py 16 0.01 > $OUT/radau_plain.log 2>&1 && head -1 $OUT/radau_plain.log &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:radau_kernel -s 1 -c 1 -o $OUT/radau_full python scripts/profile_radau.py 16 0.01 > $OUT/ncu_radau.log 2>&1; echo "ncu radau rc $?"
for t in small large; do for nb in "20000 3" "20000 4" "5000 12" "2000 32"; do
  MARLPDE_RK45_TILE=$t timeout 120 python scripts/profile_stream.py $nb 256 > $OUT/stream_tmp.log 2>&1; echo "stream $nb tile=$t: $(tail -1 $OUT/stream_tmp.log)"
done; done
echo done
