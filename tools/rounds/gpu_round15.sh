#!/bin/bash
set -u
TAG=${1:-r1v}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
echo "== bench (tf32x3 on)"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?"; tail -3 $OUT/bench.err; cut -c1-300 $OUT/bench.json
echo "== bench (tf32x3 off)"; EQF_GEMM_TF32X3=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_off.json 2> $OUT/bench_off.err; echo "rc=$?"; cut -c1-300 $OUT/bench_off.json
echo "== ncu launch list (headline region only)"
EQF_BENCH_CUDA_PROFILER=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_bench.log 2>&1 ; echo "ncu rc=$?"; wc -l $OUT/launches.csv
python tools/summarise_launches.py $OUT/launches.csv 45
