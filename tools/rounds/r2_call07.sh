#!/bin/bash
set -u
TAG=${1:-r2c07}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== kernel tests (gemm with warp-elected waits, fused)"
timeout -k 10 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x 2>&1 | tail -5
echo "== bucketed test"
timeout -k 10 600 python -m pytest tests/test_gpu_model.py -m gpu -q -x -k "bucketed" 2>&1 | tail -30
echo "== microbench qm9"
timeout -k 10 300 python tools/fused_microbench.py qm9_l2 32560 20 > $OUT/fused_microbench.jsonl 2> $OUT/mb.err; echo "rc=$?"; tail -2 $OUT/mb.err
python - <<PY
import json
for line in open("$OUT/fused_microbench.jsonl"):
    r = json.loads(line); print(r["case"], "unfused_dtp", r["unfused_dtp_us"], "fused_total", r["fused_total_us"], "unfused_total", r["unfused_total_us"])
    for g in r["groups"]: print("   ", g)
PY
timeout -k 10 200 python tools/fused_timeline.py qm9_l2 32560 2 32 1 > $OUT/timeline_l2_dtp1.txt 2> $OUT/tl.err; tail -9 $OUT/timeline_l2_dtp1.txt | cut -c1-1200
echo "== gemm microbench (r1 kernels with warp-elected waits)"
timeout -k 10 300 python tools/tf32x3_microbench.py > $OUT/tf32x3_microbench.jsonl 2>> $OUT/mb.err; tail -12 $OUT/tf32x3_microbench.jsonl | cut -c1-300
echo "== bench fused / unfused"
timeout -k 10 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_fused.json 2> $OUT/bench.err; echo "rc=$?"; tail -3 $OUT/bench.err
EQF_FUSED=0 timeout -k 10 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_unfused.json 2>> $OUT/bench.err; echo "rc=$?"
python - <<PY
import json
for n in ["fused", "unfused"]:
    try:
        d = json.load(open("$OUT/bench_%s.json" % n))
        print(n, "ms/step", d["ms_per_step"], "eager", d["config"]["eager_ms_per_step"], "e2e", d["e2e"]["ms_per_step"])
        for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["ms_per_step"])[:7]: print(f"{v['ms_per_step']:8.3f} ms {v['launches_per_step']:6.1f}  {k}")
    except Exception as e: print("no bench", e)
PY
