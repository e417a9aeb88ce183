#!/bin/bash
set -u
G=${2:-2}
OUT=gpurun_out/${1:-multi}
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 scripts/check_multi_gpu.py > $OUT/check_multi.log 2>&1
echo "check exit $?" >> $OUT/check_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $G --steps 3 --warmup 3 > $OUT/bench_g$G.json 2> $OUT/bench_g$G.err
echo "bench exit $?" >> $OUT/bench_g$G.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $G --steps 1 --warmup 1 > $OUT/bench_ref_g$G.json 2>> $OUT/bench_g$G.err
echo done
