#!/bin/bash
set -u
TAG=${1:-r1e}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== kernel tests (v3 default)"; timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_golden.py -m gpu -q -x 2>&1 | tail -6
echo "== dtp microbench"
for cfg in "v3 8" "v3 4" "tma 8"; do
  set -- $cfg
  EQF_DTP_VARIANT=$1 EQF_TILE_EDGES=$2 timeout 120 python tools/dtp_microbench.py qm9_l2 32560 20 2>&1 | tail -1 | tee -a $OUT/dtp_microbench.jsonl
done
EQF_DTP_VARIANT=v3 timeout 120 python tools/dtp_microbench.py qm9_l2 500000 5 2>&1 | tail -1 | tee -a $OUT/dtp_microbench.jsonl
EQF_DTP_VARIANT=v3 timeout 120 python tools/dtp_microbench.py md17_l3 20000 10 2>&1 | tail -1 | tee -a $OUT/dtp_microbench.jsonl
EQF_DTP_VARIANT=v3 timeout 120 python tools/dtp_microbench.py oc20_l1 58000 10 2>&1 | tail -1 | tee -a $OUT/dtp_microbench.jsonl
echo "== model tests"; timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q -x 2>&1 | tail -6
echo "== bench (graph)"; timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?"; tail -3 $OUT/bench.err; cat $OUT/bench.json
echo "== bench (eager, torch gemm)"; EQF_GEMM=torch timeout 600 python bench.py --steps 10 --warmup 3 --no-graph --no-cpu-baseline > $OUT/bench_eager_torch.json 2> $OUT/bench_eager_torch.err; echo "rc=$?"; cat $OUT/bench_eager_torch.json | python -c "import json,sys; d=json.load(sys.stdin); print(d['ms_per_step'], d['value'])"
echo "== bench (graph, torch gemm)"; EQF_GEMM=torch timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_graph_torch.json 2> $OUT/bench_graph_torch.err; echo "rc=$?"; cat $OUT/bench_graph_torch.json | python -c "import json,sys; d=json.load(sys.stdin); print(d['ms_per_step'], d['value'])"
