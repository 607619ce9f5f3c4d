#!/bin/bash
set -u
TAG=${1:-r1p}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== timeline"; timeout 120 python tools/tf32x3_timeline.py 32560 224 224
echo "== tf32x3 microbench"; timeout 180 python tools/tf32x3_microbench.py > $OUT/tf32x3.jsonl 2> $OUT/tf32x3.err; echo "rc=$?"; tail -5 $OUT/tf32x3.err
grep '"us"' $OUT/tf32x3.jsonl | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print(f\"{r['shape']:16s} M={r['M']:7d} K={r['K']:4d} N={r['N']:4d} err={r['rel_err']:.1e} us={r['us']:6.1f} cutlass={r['cutlass_us']:6.1f} cublas={r['cublas_us']:6.1f} ({r['cublas_err']:.1e}) {r['gb_s']:7.1f} GB/s {r['tflops']:5.1f} TF\")
"
