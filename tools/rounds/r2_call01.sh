#!/bin/bash
# round 2, call 1: the full GPU suite incl. the new full-size parity tests, smoke, a baseline bench line of the
# round-1 kernels on this box and a fresh launch list of the headline region (the r1 list predated the last kernel change)
set -u
TAG=${1:-r2c01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv | tail -2
free -g | head -2; nproc
echo "== pytest gpu (new full-size parity tests first)"
timeout 1500 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -x --durations=8 2>&1 | tail -25
echo "== pytest gpu (all the rest)"
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_fullsize.py --durations=5 2>&1 | tail -15
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?"; tail -3 $OUT/bench.err; cut -c1-400 $OUT/bench.json
echo "== ncu launch list (headline region only)"
EQF_BENCH_CUDA_PROFILER=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_bench.log 2>&1 ; echo "ncu rc=$?"; wc -l $OUT/launches.csv
python tools/summarise_launches.py $OUT/launches.csv 70 | tee $OUT/launches_summary.txt
