#!/bin/bash
set -u
TAG=${1:-r2c}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest tf32x3"; timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "tf32x3" 2>&1 | tail -5
echo "== fwd microbench (graph-replayed)"; timeout 240 python tools/tf32x3_microbench.py > $OUT/tf32x3.jsonl 2> $OUT/tf32x3.err; echo "rc=$?"; tail -3 $OUT/tf32x3.err
grep '"us"' $OUT/tf32x3.jsonl | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print(f\"{r['shape']:16s} M={r['M']:7d} K={r['K']:4d} N={r['N']:4d} err={r['rel_err']:.1e} us={r['us']:6.1f} cutlass={r['cutlass_us']:6.1f} cublas={r['cublas_us']:6.1f}\")
"
