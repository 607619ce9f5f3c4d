#!/bin/bash
set -u
TAG=${1:-r1v}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
echo "== tf32x3 microbench"; timeout 180 python tools/tf32x3_microbench.py > $OUT/tf32x3.jsonl 2> $OUT/tf32x3.err; echo "rc=$?"; tail -3 $OUT/tf32x3.err
grep '"us"' $OUT/tf32x3.jsonl | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print(f\"{r['shape']:16s} M={r['M']:7d} K={r['K']:4d} N={r['N']:4d} err={r['rel_err']:.1e} us={r['us']:6.1f} cutlass={r['cutlass_us']:6.1f} cublas={r['cublas_us']:6.1f}\")
"
echo "== bench (tf32x3 on)"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?"; tail -3 $OUT/bench.err; cut -c1-300 $OUT/bench.json
echo "== bench (tf32x3 off)"; EQF_GEMM_TF32X3=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_off.json 2> $OUT/bench_off.err; echo "rc=$?"; cut -c1-300 $OUT/bench_off.json
