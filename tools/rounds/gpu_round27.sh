#!/bin/bash
set -u
for skip in 0 1 2 3; do
echo "== wgrad skip=$skip"; EQF_TF32X3_DBG_SKIP=$skip timeout 120 python - <<'PY'
import sys, os, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tools')
from equiformer_b200 import ops
from tf32x3_wgrad_microbench import timeit
dev = torch.device('cuda:0')
for (R, K1, N) in [(162800, 352, 32), (97680, 384, 64), (32560, 224, 224)]:
    A = torch.randn(R, K1, device=dev); G = torch.randn(R, N, device=dev)
    us = timeit(lambda: ops.gemm_tf32x3_wgrad_raw(A, G))
    print(f"  {R}x{K1}^T x {N}: {us:.1f} us, operand stream {4*R*(K1+N)/us/1e3:.0f} GB/s")
PY
done
