#!/bin/bash
# r02J: cost of a rendezvous of the two CTAs of a cluster (scripts/ubench/cluster_sync.cu)
set -u
OUT=gpurun_out/${1:-r02J}; mkdir -p $OUT
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/cluster_sync scripts/ubench/cluster_sync.cu > $OUT/build.log 2>&1
timeout 60 /tmp/cluster_sync > $OUT/cluster_sync.log 2>&1; echo "exit $?" >> $OUT/cluster_sync.log
cat $OUT/cluster_sync.log
echo done
