#!/bin/bash
# GEMM band / promotion-interval variants + fused LayerNorm tests + bench
set -u
TAG=${1:-r1l}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest eln"; timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "layer_norm" 2>&1 | tail -4
for v in b3p1k16 b3p1k32 b3p2k32 b3p4k64 b4p1k16 b5p2k32; do
  if [ -f equiformer_b200/libeqf_gemm_$v.so ]; then
    echo "== gemm variant $v"
    EQF_GEMM_LIB=$PWD/equiformer_b200/libeqf_gemm_$v.so EQF_GEMM_FORCE=1 timeout 300 python tools/gemm_microbench.py > $OUT/gemm_$v.jsonl 2> $OUT/gemm_$v.err; echo "rc=$?"; tail -2 $OUT/gemm_$v.err
    python - <<PY
import json
tot=0; worst=0
for l in open("$OUT/gemm_$v.jsonl"):
    r=json.loads(l)
    for m in ("mode0","mode1","mode2"):
        if m in r: tot+=r[m]["us"]; worst=max(worst,r[m]["rel_err"])
print("$v total_us=%.1f worst_rel_err=%.2e"%(tot,worst))
PY
  fi
done
echo "== stock"; EQF_GEMM_FORCE=1 timeout 300 python tools/gemm_microbench.py > $OUT/gemm_stock.jsonl 2>/dev/null
python - <<PY
import json
tot=0; worst=0
for l in open("$OUT/gemm_stock.jsonl"):
    r=json.loads(l)
    for m in ("mode0","mode1","mode2"):
        if m in r: tot+=r[m]["us"]; worst=max(worst,r[m]["rel_err"])
print("stock total_us=%.1f worst_rel_err=%.2e"%(tot,worst))
PY
echo "== bench (graph, fused LN)"; timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?"; tail -3 $OUT/bench.err; cut -c1-400 $OUT/bench.json
