#!/bin/bash
set -u
TAG=${1:-r2c10}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== previously failing + new tests"
timeout -k 10 1500 python -m pytest tests -m gpu -q -x -k "fixture or model_file or edge_geometry or expnorm or attention_family or tcgen05 or eager or md17" --durations=3 2>&1 | tail -12
echo "== bench qm9 / md17"
timeout -k 10 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?"; tail -3 $OUT/bench.err
timeout -k 10 900 python bench.py --workload md17_l3 --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_md17_l3.json 2> $OUT/bench_md17.err; echo "rc=$?"; tail -3 $OUT/bench_md17.err
python - <<PY
import json
for n in ["bench", "bench_md17_l3"]:
    d = json.load(open("$OUT/%s.json" % n))
    print(n, "ms/step", d["ms_per_step"], "value", d["value"], "launches", d["gpu_launches"], "eager", d["config"]["eager_ms_per_step"])
PY
