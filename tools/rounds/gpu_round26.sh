#!/bin/bash
set -u
OUT=gpurun_out/${1:-r2s}; mkdir -p $OUT
echo "== pytest tf32x3"; timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "tf32x3" 2>&1 | tail -3
for sets in 2 1; do
echo "== fwd microbench SETS=$sets"; EQF_TF32X3_SETS=$sets timeout 240 python tools/tf32x3_microbench.py > $OUT/tf32x3_$sets.jsonl 2> $OUT/tf32x3_$sets.err; tail -2 $OUT/tf32x3_$sets.err
grep '"us"' $OUT/tf32x3_$sets.jsonl | python -c "
import sys,json
print(' '.join(f\"{json.loads(l)['shape']}={json.loads(l)['us']:.1f}({json.loads(l)['rel_err']:.0e})\" for l in sys.stdin))"
done
