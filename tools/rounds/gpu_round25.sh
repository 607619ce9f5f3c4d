#!/bin/bash
set -u
TAG=${1:-r2r}
OUT=gpurun_out/$TAG
mkdir -p $OUT
cat > /tmp/one_gemm.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
from equiformer_b200 import ops
dev = torch.device('cuda:0')
M, K, N = (int(v) for v in sys.argv[1:4])
A = torch.randn(M, K, device=dev); Bt = torch.randn(N, K, device=dev)
for _ in range(5):
    ops.gemm_tf32x3_raw(A, Bt)
torch.cuda.synchronize()
PY
echo "== ncu full: wide kernel (224 -> 224)"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3_kernel -s 3 -c 1 -o $OUT/prof_tf32x3_wide python /tmp/one_gemm.py 32560 224 224 > $OUT/ncu_wide.log 2>&1; echo "rc=$?"
echo "== ncu full: TMEM-A stacked kernel (352 -> 32)"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3_ts_kernel -s 3 -c 1 -o $OUT/prof_tf32x3_ts python /tmp/one_gemm.py 162800 352 32 > $OUT/ncu_ts.log 2>&1; echo "rc=$?"
for n in wide ts; do ncu -i $OUT/prof_tf32x3_$n.ncu-rep --page raw --csv > $OUT/prof_tf32x3_${n}_raw.csv 2>/dev/null; wc -l $OUT/prof_tf32x3_${n}_raw.csv; done
