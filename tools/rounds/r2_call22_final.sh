#!/bin/bash
# final validation of the round: full GPU suite, smoke, default bench (with the CPU baseline), reference arms, other workloads
set -u
TAG=${1:-r2c22}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== full gpu suite"
timeout -k 10 1800 python -m pytest tests -m gpu -q --durations=3 2>&1 | tail -8
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== bench default"
timeout -k 10 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?"; tail -2 $OUT/bench.err
python - <<PY
import json
d = json.load(open("$OUT/bench.json"))
print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"], "cpu", d["cpu_baseline"])
print("roof", {k: d["roofline"].get(k) for k in ("kernel", "bound", "frac", "hbm", "tensor", "traffic")})
for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["ms_per_step"])[:14]: print(f"{v['ms_per_step']:8.3f} ms {v['launches_per_step']:6.1f}  {k}")
PY
echo "== other workloads"
for w in md17_l3 oc20_l1 stress; do
  timeout -k 10 900 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_$w.json 2> $OUT/bench_$w.err; echo "$w rc=$?"
done
timeout -k 10 600 python bench.py --stream 16 --steps 16 --warmup 3 --no-cpu-baseline > $OUT/bench_qm9_stream.json 2> $OUT/bench_stream.err; echo "stream rc=$?"
python - <<PY
import json
for n in ["bench_md17_l3", "bench_oc20_l1", "bench_stress", "bench_qm9_stream"]:
    try:
        d = json.load(open("$OUT/%s.json" % n))
        print(n, "ms/step", round(d["ms_per_step"], 3), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"], "roofline", d["roofline"].get("frac"), d["roofline"].get("kernel"))
    except Exception as e:
        print(n, "failed", e)
PY
echo "== reference arms"
timeout -k 10 900 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "rc=$?"; cut -c1-500 $OUT/bench_ref.json
timeout -k 10 900 python bench.py --impl reference-gpu --steps 3 --warmup 1 > $OUT/bench_refgpu.json 2> $OUT/bench_refgpu.err; echo "rc=$?"; cut -c1-400 $OUT/bench_refgpu.json
