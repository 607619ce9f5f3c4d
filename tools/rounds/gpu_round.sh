#!/bin/bash
# One GPU-box round: smoke, GPU parity tests, bench, ncu launch list.  Outputs under gpurun_out/.
# usage: tools/gpu_round.sh [tag]
set -u
TAG=${1:-r1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.csv 2>&1
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $OUT/smoke.log 2>&1 ; echo "smoke rc=$?" ; tail -3 $OUT/smoke.log
echo "== pytest gpu" ; timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > $OUT/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -15 $OUT/pytest_gpu.log
echo "== bench" ; timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err ; echo "bench rc=$?" ; tail -2 $OUT/bench.err ; cat $OUT/bench.json
if [ "${VARIANTS:-0}" = "1" ]; then
for v in scalar vec; do
  echo "== kernels with EQF_DTP_VARIANT=$v" ; EQF_DTP_VARIANT=$v timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "dtp" > $OUT/pytest_$v.log 2>&1 ; echo "rc=$?" ; tail -3 $OUT/pytest_$v.log
done
for cfg in "vec cutlass" "scalar cutlass" "tma torch"; do
  set -- $cfg
  echo "== bench DTP=$1 GEMM=$2" ; EQF_DTP_VARIANT=$1 EQF_GEMM=$2 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_$1_$2.json 2> $OUT/bench_$1_$2.err ; echo "rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_$1_$2.json"))
    print("  ms/step", round(d["ms_per_step"],2), "edges/s", int(d["value"]), {k:(round(v["ms_per_step"],3), v["gb_s"] and int(v["gb_s"])) for k,v in d["kernels"].items()})
except Exception as e: print("  parse failed", e)
PY
done
fi
if [ "${NCU:-1}" = "1" ]; then
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_bench.log 2>&1 ; echo "ncu rc=$?"
fi
echo done
