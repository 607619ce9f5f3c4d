#!/bin/bash
# 8 GPUs: weak scaling of the OC20 frames and the stress cell (one independent batch / cell per rank), and the headline
set -u
TAG=${1:-r2c11}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi -L | head -8
for WL in oc20_l1 stress qm9; do
  for N in 8; do
    timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $WL --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_${WL}_n$N.json 2> $OUT/bench_${WL}_n$N.err
    echo "$WL N=$N rc=$?"; tail -2 $OUT/bench_${WL}_n$N.err | cut -c1-200; cut -c1-240 $OUT/bench_${WL}_n$N.json
  done
done
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload oc20_l1 --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_oc20_l1_n2.json 2> $OUT/bench_oc20_l1_n2.err; echo "oc20 N=2 rc=$?"; cut -c1-240 $OUT/bench_oc20_l1_n2.json
