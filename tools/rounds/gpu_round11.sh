#!/bin/bash
# colsum / add_bias / fused FFN gate: all GPU tests, bench (default and node-level tcgen05 wgrad), per-shape torch profile
set -u
TAG=${1:-r1n}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?"; tail -3 $OUT/bench.err; cut -c1-300 $OUT/bench.json
echo "== bench EQF_WGRAD_MIN_K=2048"; EQF_WGRAD_MIN_K=2048 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_wgrad2048.json 2> $OUT/bench_wgrad2048.err; echo "rc=$?"; tail -3 $OUT/bench_wgrad2048.err; cut -c1-300 $OUT/bench_wgrad2048.json
echo "== profile step"; timeout 300 python tools/profile_step.py $OUT 2>&1 | head -40 | cut -c1-160
echo "== shapes"; head -70 $OUT/profile_step_shapes.txt | cut -c1-230
