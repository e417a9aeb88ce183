#!/bin/bash
# N-GPU bench line only (the driver's SCALE command), on the final tree
set -u
G=${2:-2}
OUT=gpurun_out/${1:-multi}
mkdir -p $OUT
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $G --steps 3 --warmup 3 ) > $OUT/bench_g$G.json 2> $OUT/bench_g$G.err
echo "bench exit $?" >> $OUT/bench_g$G.err
tail -5 $OUT/bench_g$G.err; cut -c1-400 $OUT/bench_g$G.json
echo done
