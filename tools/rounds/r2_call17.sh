#!/bin/bash
# grouped small-product kernel: parity, model-level fixtures, headline / MD17 A/B on one box
set -u
TAG=${1:-r2c17}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== new tests"
timeout -k 10 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "grouped" 2>&1 | tail -8
echo "== model-level tests"
timeout -k 10 1200 python -m pytest tests/test_reference_golden.py tests/test_gpu_model.py -m gpu -q -x 2>&1 | tail -5
echo "== bench qm9 grouped on / off"
timeout -k 10 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_qm9.json 2> $OUT/bench_qm9.err; echo "rc=$?"; tail -2 $OUT/bench_qm9.err
EQF_GROUPED_GEMM=0 timeout -k 10 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_qm9_ungrouped.json 2> $OUT/bench_qm9_u.err; echo "rc=$?"
echo "== bench md17 grouped on / off"
timeout -k 10 900 python bench.py --workload md17_l3 --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_md17.json 2> $OUT/bench_md17.err; echo "rc=$?"; tail -2 $OUT/bench_md17.err
EQF_GROUPED_GEMM=0 timeout -k 10 900 python bench.py --workload md17_l3 --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_md17_ungrouped.json 2> $OUT/bench_md17_u.err; echo "rc=$?"
python - <<PY
import json
for n in ["bench_qm9", "bench_qm9_ungrouped", "bench_md17", "bench_md17_ungrouped"]:
    try:
        d = json.load(open("$OUT/%s.json" % n))
        print(n, "ms/step", round(d["ms_per_step"], 3), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"], "roofline", d["roofline"].get("frac"), d["roofline"].get("kernel"))
    except Exception as e:
        print(n, "failed", e)
PY
