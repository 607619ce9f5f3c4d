#!/bin/bash
# round 2, call 2: first run of K1-forward (eqf_dtp_linear_fwd): kernel parity, micro-benchmark with the skip-mode
# diagnostics, model-level parity with the fused path on, bench line
set -u
TAG=${1:-r2c02}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== fused kernel parity"
timeout -k 10 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "fused" 2>&1 | tail -15
echo "== microbench qm9"
timeout -k 10 300 python tools/fused_microbench.py qm9_l2 32560 20 > $OUT/fused_microbench.jsonl 2> $OUT/mb.err; echo "rc=$?"; tail -2 $OUT/mb.err; cat $OUT/fused_microbench.jsonl
for SK in 1 2 4 3; do
  echo "== microbench skip=$SK"
  EQF_FUSED_DBG_SKIP=$SK timeout -k 10 300 python tools/fused_microbench.py qm9_l2 32560 10 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    r = json.loads(line); print(r['case'], [(g['l'], g['N'], g['fused_us']) for g in r['groups']])"
done | tee $OUT/fused_skip_modes.txt
echo "== microbench oc20 / md17"
timeout -k 10 300 python tools/fused_microbench.py oc20_l1 58000 10 >> $OUT/fused_microbench.jsonl 2>> $OUT/mb.err; tail -2 $OUT/fused_microbench.jsonl | cut -c1-600
timeout -k 10 300 python tools/fused_microbench.py md17_l3 1700 10 >> $OUT/fused_microbench.jsonl 2>> $OUT/mb.err; tail -2 $OUT/fused_microbench.jsonl | cut -c1-600
echo "== model-level parity (fused on)"
timeout -k 10 1500 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_model.py tests/test_reference_golden.py -m gpu -q --durations=5 2>&1 | tail -15
echo "== bench fused"; timeout -k 10 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_fused.json 2> $OUT/bench.err; echo "rc=$?"; tail -3 $OUT/bench.err; cut -c1-300 $OUT/bench_fused.json
python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_fused.json"))
    print("ms/step", d["ms_per_step"], "eager", d["config"]["eager_ms_per_step"], "e2e", d["e2e"]["ms_per_step"])
    for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["ms_per_step"]): print(f"{v['ms_per_step']:8.3f} ms {v['launches_per_step']:6.1f}  {k}")
except Exception as e: print("no bench", e)
PY
echo "== bench unfused (same box)"; EQF_FUSED=0 timeout -k 10 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_unfused.json 2>> $OUT/bench.err; echo "rc=$?"; cut -c1-200 $OUT/bench_unfused.json
