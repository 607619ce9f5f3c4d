#!/bin/bash
set -u
TAG=${1:-r1h}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== gemm tests"; timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "gemm" 2>&1 | tail -4
echo "== gemm microbench (forced cutlass, 1SM)"; EQF_GEMM_FORCE=1 timeout 300 python tools/gemm_microbench.py 32560 2>&1 | tee $OUT/gemm_microbench_1sm.jsonl | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print(d['shape'], {k:(v['us'], v['cublas_us'], v['rel_err']) for k,v in d.items() if k.startswith('mode')})"
echo "== gemm microbench (forced cutlass, 2SM)"; EQF_GEMM_2SM=1 EQF_GEMM_FORCE=1 timeout 300 python tools/gemm_microbench.py 32560 2>&1 | tee $OUT/gemm_microbench_2sm.jsonl | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print(d['shape'], {k:(v['us'], v['cublas_us'], v['rel_err']) for k,v in d.items() if k.startswith('mode')})"
echo "== pytest gpu (model)"; timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q -x 2>&1 | tail -4
echo "== bench (graph)"; timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?"; tail -3 $OUT/bench.err; cat $OUT/bench.json
echo "== profile step"; timeout 300 python tools/profile_step.py $OUT 2>&1 | tail -30
