#!/bin/bash
set -u
TAG=${1:-r2c08}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== fused parity + bucketed"
timeout -k 10 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q -x -k "fused or bucketed" 2>&1 | tail -12
echo "== microbench qm9 (per-degree kernels: smaller code)"
timeout -k 10 300 python tools/fused_microbench.py qm9_l2 32560 20 > $OUT/fused_microbench.jsonl 2> $OUT/mb.err; echo "rc=$?"; tail -2 $OUT/mb.err
python - <<PY
import json
for line in open("$OUT/fused_microbench.jsonl"):
    r = json.loads(line); print(r["case"], "unfused_dtp", r["unfused_dtp_us"], "fused_total", r["fused_total_us"], "unfused_total", r["unfused_total_us"])
    for g in r["groups"]: print("   ", g)
PY
for SK in 1 7; do
  echo "== microbench skip=$SK"
  EQF_FUSED_DBG_SKIP=$SK timeout -k 10 300 python tools/fused_microbench.py qm9_l2 32560 10 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    r = json.loads(line); print(r['case'], [(g['l'], g['N'], g['fused_us']) for g in r['groups']])"
done | tee $OUT/fused_skip_modes.txt
echo "== bench md17 (graph) / stress / oc20"
for WL in md17_l3 stress; do
  timeout -k 10 1200 python bench.py --workload $WL --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_$WL.json 2> $OUT/bench_$WL.err; echo "$WL rc=$?"; tail -3 $OUT/bench_$WL.err; cut -c1-260 $OUT/bench_$WL.json
done
EQF_FUSED=0 timeout -k 10 1200 python bench.py --workload stress --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_stress_unfused.json 2> $OUT/bench_stress_unfused.err; echo "stress unfused rc=$?"; tail -3 $OUT/bench_stress_unfused.err; cut -c1-260 $OUT/bench_stress_unfused.json
EQF_FUSED=0 timeout -k 10 600 python bench.py --workload oc20_l1 --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_oc20_unfused.json 2> $OUT/bench_oc20_unfused.err; echo "oc20 unfused rc=$?"; cut -c1-260 $OUT/bench_oc20_unfused.json
echo "== reference-gpu arm"
timeout -k 10 600 python bench.py --impl reference-gpu --steps 3 --warmup 1 > $OUT/bench_refgpu.json 2> $OUT/bench_refgpu.err; echo "rc=$?"; tail -3 $OUT/bench_refgpu.err; cut -c1-300 $OUT/bench_refgpu.json
