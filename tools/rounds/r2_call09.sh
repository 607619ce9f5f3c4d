#!/bin/bash
set -u
TAG=${1:-r2c09}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== full gpu suite (auto policy)"
timeout -k 10 2400 python -m pytest tests -m gpu -q --durations=5 2>&1 | tail -12
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== bench default"
timeout -k 10 900 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?"; tail -3 $OUT/bench.err
python - <<PY
import json
d = json.load(open("$OUT/bench.json"))
print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"], "launches", d["gpu_launches"], "cpu", d["cpu_baseline"])
print("roof", {k: d["roofline"][k] for k in ("kernel", "bound", "frac", "hbm", "tensor")})
for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["ms_per_step"])[:12]: print(f"{v['ms_per_step']:8.3f} ms {v['launches_per_step']:6.1f}  {k}")
PY
echo "== reference arm (CPU, 32 molecules, one socket)"
timeout -k 10 900 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "rc=$?"; cut -c1-700 $OUT/bench_ref.json
echo "== ncu launch list (headline region only)"
EQF_BENCH_CUDA_PROFILER=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_bench.log 2>&1 ; echo "ncu rc=$?"; wc -l $OUT/launches.csv
python tools/summarise_launches.py $OUT/launches.csv 45 | tee $OUT/launches_summary.txt
