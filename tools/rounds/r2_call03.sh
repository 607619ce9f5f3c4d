#!/bin/bash
# round 2, call 3: K1-forward v2 (table helper warps, TMA radial-weight box, structural-zero mask, wide groups to HBM)
set -u
TAG=${1:-r2c03}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== fused kernel parity"
timeout -k 10 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "fused" 2>&1 | tail -8
echo "== microbench qm9"
timeout -k 10 300 python tools/fused_microbench.py qm9_l2 32560 20 > $OUT/fused_microbench.jsonl 2> $OUT/mb.err; echo "rc=$?"; tail -2 $OUT/mb.err
python - <<PY
import json
for line in open("$OUT/fused_microbench.jsonl"):
    r = json.loads(line); print(r["case"], "unfused_dtp", r["unfused_dtp_us"], "fused_total", r["fused_total_us"], "unfused_total", r["unfused_total_us"])
    for g in r["groups"]: print("   ", g)
PY
for SK in 1 2 4 3; do
  echo "== microbench skip=$SK"
  EQF_FUSED_DBG_SKIP=$SK timeout -k 10 300 python tools/fused_microbench.py qm9_l2 32560 10 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    r = json.loads(line); print(r['case'], [(g['l'], g['N'], g['fused_us']) for g in r['groups']])"
done | tee $OUT/fused_skip_modes.txt
echo "== microbench oc20 / md17 / stress"
for C in "oc20_l1 58000" "md17_l3 1700" "qm9_l2 500000"; do
timeout -k 10 300 python tools/fused_microbench.py $C 10 2>> $OUT/mb.err | tee -a $OUT/fused_microbench.jsonl | python -c "
import sys, json
for line in sys.stdin:
    r = json.loads(line); print(r['config'], r['E'], r['case'], 'fused', r['fused_total_us'], 'unfused', r['unfused_total_us'], [(g['l'], g['N'], g['gemm_us'], g['fused_us'], g.get('route_us')) for g in r['groups']])"
done
echo "== model-level parity (fused on)"
timeout -k 10 1500 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_model.py tests/test_reference_golden.py -m gpu -q --durations=3 2>&1 | tail -12
echo "== bench fused"; timeout -k 10 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_fused.json 2> $OUT/bench.err; echo "rc=$?"; tail -3 $OUT/bench.err
python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_fused.json"))
    print("ms/step", d["ms_per_step"], "eager", d["config"]["eager_ms_per_step"], "e2e", d["e2e"]["ms_per_step"])
    for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["ms_per_step"]): print(f"{v['ms_per_step']:8.3f} ms {v['launches_per_step']:6.1f}  {k}")
except Exception as e: print("no bench", e)
PY
