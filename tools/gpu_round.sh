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
if [ "${NCU:-1}" = "1" ]; then
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'dtp_|seg_softmax|aggregate|edge_dot|edge_scale' -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_bench.log 2>&1 ; echo "ncu rc=$?"
fi
echo done
