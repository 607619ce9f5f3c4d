#!/bin/bash
set -u
TAG=${1:-r2o}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest tf32x3"; timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "tf32x3" 2>&1 | tail -5
echo "== fwd microbench"; timeout 240 python tools/tf32x3_microbench.py > $OUT/tf32x3.jsonl 2> $OUT/tf32x3.err; tail -2 $OUT/tf32x3.err
grep '"us"' $OUT/tf32x3.jsonl | python -c "
import sys,json
print(' '.join(f\"{json.loads(l)['shape']}={json.loads(l)['us']:.1f}({json.loads(l)['rel_err']:.0e})\" for l in sys.stdin))"
echo "== wgrad microbench"; timeout 240 python tools/tf32x3_wgrad_microbench.py > $OUT/wgrad.jsonl 2> $OUT/wgrad.err; tail -2 $OUT/wgrad.err
grep '"us"' $OUT/wgrad.jsonl | python -c "
import sys,json
print(' '.join(f\"{json.loads(l)['shape']}={json.loads(l)['us']:.1f}({json.loads(l)['rel_err']:.0e})\" for l in sys.stdin))"
