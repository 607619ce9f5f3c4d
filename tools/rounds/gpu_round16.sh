#!/bin/bash
set -u
TAG=${1:-r1w}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== wgrad microbench (graph-replayed)"; timeout 240 python tools/tf32x3_wgrad_microbench.py > $OUT/wgrad.jsonl 2> $OUT/wgrad.err; echo "rc=$?"; tail -5 $OUT/wgrad.err
grep '"us"' $OUT/wgrad.jsonl | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print(f\"{r['shape']:12s} R={r['R']:7d} K1={r['K1']:4d} N={r['N']:4d} err={r['rel_err']:.1e} us={r['us']:6.1f} cutlass={r['cutlass_sliced_us']:6.1f} cublas={r['cublas_us']:6.1f} {r['gb_s']:7.1f} GB/s\")
"
echo "== fwd microbench (graph-replayed)"; timeout 240 python tools/tf32x3_microbench.py > $OUT/tf32x3.jsonl 2> $OUT/tf32x3.err; echo "rc=$?"; tail -3 $OUT/tf32x3.err
grep '"us"' $OUT/tf32x3.jsonl | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print(f\"{r['shape']:16s} M={r['M']:7d} K={r['K']:4d} N={r['N']:4d} err={r['rel_err']:.1e} us={r['us']:6.1f} cutlass={r['cutlass_us']:6.1f} cublas={r['cublas_us']:6.1f}\")
"
