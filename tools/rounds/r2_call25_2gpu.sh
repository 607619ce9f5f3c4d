#!/bin/bash
# 2 GPUs: the final code through torchrun / NCCL (headline, OC20) incl. the reference arm's rank handling
set -u
TAG=${1:-r2c25}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for WL in qm9 oc20_l1; do
  timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload $WL --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_${WL}_n2.json 2> $OUT/bench_${WL}_n2.err
  echo "$WL N=2 rc=$?"; tail -1 $OUT/bench_${WL}_n2.err | cut -c1-200; cut -c1-260 $OUT/bench_${WL}_n2.json
done
