#!/bin/bash
# 8 GPUs, final code: headline and OC20 (weak scaling, one batch per rank, one NCCL all-reduce of the flat gradient bucket)
set -u
TAG=${1:-r2c26}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for WL in qm9 oc20_l1; do
  timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --workload $WL --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_${WL}_n8.json 2> $OUT/bench_${WL}_n8.err
  echo "$WL N=8 rc=$?"; cut -c1-260 $OUT/bench_${WL}_n8.json
done
