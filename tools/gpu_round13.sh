#!/bin/bash
set -u
TAG=${1:-r1p}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== tf32x3 microbench"; timeout 180 python tools/tf32x3_microbench.py > $OUT/tf32x3.jsonl 2> $OUT/tf32x3.err; echo "rc=$?"; tail -5 $OUT/tf32x3.err; cat $OUT/tf32x3.jsonl
nvidia-smi --query-gpu=name,memory.used --format=csv
