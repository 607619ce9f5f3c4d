#!/bin/bash
set -u
TAG=${1:-r2c06}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout -k 10 200 python tools/fused_timeline.py qm9_l2 32560 2 32 1 > $OUT/timeline_l2_dtp1.txt 2> $OUT/tl.err; echo rc=$?; tail -3 $OUT/tl.err
timeout -k 10 200 python tools/fused_timeline.py qm9_l2 32560 2 32 2 > $OUT/timeline_l2_dtp2.txt 2>> $OUT/tl.err
timeout -k 10 200 python tools/fused_timeline.py qm9_l2 32560 1 64 1 > $OUT/timeline_l1_dtp1.txt 2>> $OUT/tl.err
EQF_FUSED_DBG_SKIP=7 timeout -k 10 200 python tools/fused_timeline.py qm9_l2 32560 2 32 1 > $OUT/timeline_l2_dtp1_skip7.txt 2>> $OUT/tl.err
tail -12 $OUT/timeline_l2_dtp1.txt | cut -c1-1500
echo ---- skip7
tail -12 $OUT/timeline_l2_dtp1_skip7.txt | cut -c1-1500
echo "== new tests (bucketed stream, small-model tcgen05 grads)"
timeout -k 10 900 python -m pytest tests/test_gpu_model.py tests/test_reference_golden.py -m gpu -q -x -k "bucketed or tcgen05" 2>&1 | tail -6
echo "== bench lines of the new bench.py"
for WL in qm9 oc20_l1 md17_l3; do
  timeout -k 10 900 python bench.py --workload $WL --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_$WL.json 2> $OUT/bench_$WL.err; echo "$WL rc=$?"; tail -2 $OUT/bench_$WL.err; cut -c1-330 $OUT/bench_$WL.json
done
timeout -k 10 900 python bench.py --workload qm9 --stream 16 --steps 16 --warmup 3 --no-cpu-baseline > $OUT/bench_qm9_stream.json 2> $OUT/bench_qm9_stream.err; echo "stream rc=$?"; tail -2 $OUT/bench_qm9_stream.err
python - <<PY
import json
for n in ["qm9", "oc20_l1", "md17_l3", "qm9_stream"]:
    try:
        d = json.load(open("$OUT/bench_%s.json" % n))
        print(n, "ms/step", round(d["ms_per_step"], 3), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["e2e"].get("captures"), "eager", round(d["config"]["eager_ms_per_step"], 2), "launches", d["gpu_launches"])
        print("   roof", d["roofline"]["kernel"], d["roofline"]["bound"], round(d["roofline"]["frac"], 3), "hbm", round(d["roofline"]["hbm"]["frac"], 3), "tensor", round(d["roofline"]["tensor"]["frac"], 3))
    except Exception as e: print(n, "failed", e)
PY
