#!/bin/bash
set -u
TAG=${1:-r1q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
cat > /tmp/one_gemm.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
from equiformer_b200 import ops
dev = torch.device('cuda:0')
M, K, N = 32560, 224, 224
A = torch.randn(M, K, device=dev); Bt = torch.randn(N, K, device=dev)
for _ in range(5):
    ops.gemm_tf32x3_raw(A, Bt)
torch.cuda.synchronize()
PY
echo "== ncu full: tf32x3 val1_l0"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3_kernel -s 3 -c 1 -o $OUT/prof_tf32x3 python /tmp/one_gemm.py > $OUT/ncu_tf32x3.log 2>&1; echo "rc=$?"; tail -3 $OUT/ncu_tf32x3.log
ncu -i $OUT/prof_tf32x3.ncu-rep --page raw --csv > $OUT/prof_tf32x3_raw.csv 2>/dev/null
ncu -i $OUT/prof_tf32x3.ncu-rep --page source --csv > $OUT/prof_tf32x3_source.csv 2>/dev/null
wc -l $OUT/prof_tf32x3_raw.csv $OUT/prof_tf32x3_source.csv
echo "== pytest graph/model"; timeout 900 python -m pytest tests -m gpu -q -x -k "graph or model" 2>&1 | tail -4
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?"; tail -3 $OUT/bench.err; cut -c1-300 $OUT/bench.json
