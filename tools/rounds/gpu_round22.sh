#!/bin/bash
set -u
TAG=${1:-r2k}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for promo in 256 128; do
echo "== L2 promotion $promo: fwd"; EQF_TF32X3_L2PROMO=$promo timeout 240 python tools/tf32x3_microbench.py > $OUT/tf32x3_$promo.jsonl 2> $OUT/tf32x3_$promo.err; tail -2 $OUT/tf32x3_$promo.err
grep '"us"' $OUT/tf32x3_$promo.jsonl | python -c "
import sys,json
print(' '.join(f\"{json.loads(l)['shape']}={json.loads(l)['us']:.1f}\" for l in sys.stdin))"
echo "== L2 promotion $promo: wgrad"; EQF_TF32X3_L2PROMO=$promo timeout 240 python tools/tf32x3_wgrad_microbench.py > $OUT/wgrad_$promo.jsonl 2> $OUT/wgrad_$promo.err; tail -2 $OUT/wgrad_$promo.err
grep '"us"' $OUT/wgrad_$promo.jsonl | python -c "
import sys,json
print(' '.join(f\"{json.loads(l)['shape']}={json.loads(l)['us']:.1f}\" for l in sys.stdin))"
done
