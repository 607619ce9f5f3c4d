#!/bin/bash
set -u
TAG=${1:-r1k}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest gpu (all)"; timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
echo "== bench (graph)"; timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?"; tail -3 $OUT/bench.err; cat $OUT/bench.json
echo "== profile step"; timeout 300 python tools/profile_step.py $OUT 2>&1 | tail -42
