#!/bin/bash
# after the wait-hint / fast-rounding defaults and K1 v5: full GPU suite, headline bench (A/B against the old waits), stress, oc20
set -u
TAG=${1:-r2c16}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== full gpu suite"
timeout -k 10 1500 python -m pytest tests -m gpu -q -x --durations=3 2>&1 | tail -8
echo "== bench qm9 (new defaults), then with the hint and the fast rounding off (same box)"
timeout -k 10 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_qm9.json 2> $OUT/bench_qm9.err; echo "rc=$?"; tail -2 $OUT/bench_qm9.err
EQF_TF32X3_DBG_SKIP=24 timeout -k 10 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_qm9_oldwaits.json 2> $OUT/bench_qm9_old.err; echo "rc=$?"
EQF_FUSED=1 timeout -k 10 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_qm9_fused_v5.json 2> $OUT/bench_qm9_f.err; echo "rc=$?"
echo "== stress (K1 v5), oc20"
timeout -k 10 900 python bench.py --workload stress --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_stress.json 2> $OUT/bench_stress.err; echo "rc=$?"; tail -2 $OUT/bench_stress.err
timeout -k 10 600 python bench.py --workload oc20_l1 --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_oc20.json 2> $OUT/bench_oc20.err; echo "rc=$?"
python - <<PY
import json
for n in ["bench_qm9", "bench_qm9_oldwaits", "bench_qm9_fused_v5", "bench_stress", "bench_oc20"]:
    try:
        d = json.load(open("$OUT/%s.json" % n))
        print(n, "ms/step", round(d["ms_per_step"], 3), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"], "roofline", d["roofline"].get("frac"), d["roofline"].get("kernel"))
    except Exception as e:
        print(n, "failed", e)
PY
