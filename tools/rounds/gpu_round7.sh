#!/bin/bash
set -u
TAG=${1:-r1i}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== kernel tests"; timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x 2>&1 | tail -4
for w in 1 2 4; do
echo "== wgrad microbench waves=$w"; ONLY_WGRAD=1 EQF_WGRAD_WAVES=$w EQF_GEMM_FORCE=1 timeout 300 python tools/gemm_microbench.py 32560 2>&1 | tee $OUT/gemm_wgrad_w$w.jsonl | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print(d['shape'], {k:(v['us'], v['cublas_us'], v['rel_err']) for k,v in d.items() if k.startswith('mode')})"
done
echo "== dtp microbench (gy)"; timeout 120 python tools/dtp_microbench.py qm9_l2 32560 20 2>&1 | tail -1 | tee -a $OUT/dtp_microbench.jsonl
echo "== pytest gpu (model+golden)"; timeout 900 python -m pytest tests/test_gpu_model.py tests/test_golden.py -m gpu -q -x 2>&1 | tail -4
echo "== bench (graph)"; timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?"; tail -3 $OUT/bench.err; cat $OUT/bench.json
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "rc=$?"; cat $OUT/bench_ref.json | cut -c1-400
