#!/bin/bash
set -u
TAG=${1:-r2e}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest tf32x3"; timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "tf32x3" 2>&1 | tail -3
echo "== wgrad microbench BKR=32 (graph-replayed)"; timeout 240 python tools/tf32x3_wgrad_microbench.py > $OUT/wgrad.jsonl 2> $OUT/wgrad.err; echo "rc=$?"; tail -5 $OUT/wgrad.err
grep '"us"' $OUT/wgrad.jsonl | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print(f\"{r['shape']:12s} R={r['R']:7d} K1={r['K1']:4d} N={r['N']:4d} err={r['rel_err']:.1e} us={r['us']:6.1f} cutlass={r['cutlass_sliced_us']:6.1f} cublas={r['cublas_us']:6.1f} {r['gb_s']:7.1f} GB/s\")
"
