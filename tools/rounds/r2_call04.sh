#!/bin/bash
# round 2, call 4: K1-forward v2 with the parallel table helper; OC20 mirror + PBC neighbour list on the GPU
set -u
TAG=${1:-r2c04}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== fused kernel parity + oc20 model + pbc"
timeout -k 10 600 python -m pytest tests/test_gpu_kernels.py tests/test_reference_golden.py -m gpu -q -x -k "fused or oc20" 2>&1 | tail -8
echo "== microbench qm9"
timeout -k 10 300 python tools/fused_microbench.py qm9_l2 32560 20 > $OUT/fused_microbench.jsonl 2> $OUT/mb.err; echo "rc=$?"; tail -2 $OUT/mb.err
python - <<PY
import json
for line in open("$OUT/fused_microbench.jsonl"):
    r = json.loads(line); print(r["case"], "unfused_dtp", r["unfused_dtp_us"], "fused_total", r["fused_total_us"], "unfused_total", r["unfused_total_us"])
    for g in r["groups"]: print("   ", g)
PY
for SK in 1 2 4 3 7; do
  echo "== microbench skip=$SK"
  EQF_FUSED_DBG_SKIP=$SK timeout -k 10 300 python tools/fused_microbench.py qm9_l2 32560 10 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    r = json.loads(line); print(r['case'], [(g['l'], g['N'], g['fused_us']) for g in r['groups']])"
done | tee $OUT/fused_skip_modes.txt
echo "== bench fused"; timeout -k 10 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_fused.json 2> $OUT/bench.err; echo "rc=$?"; tail -3 $OUT/bench.err
python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_fused.json"))
    print("ms/step", d["ms_per_step"], "eager", d["config"]["eager_ms_per_step"], "e2e", d["e2e"]["ms_per_step"])
    for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["ms_per_step"])[:8]: print(f"{v['ms_per_step']:8.3f} ms {v['launches_per_step']:6.1f}  {k}")
except Exception as e: print("no bench", e)
PY
