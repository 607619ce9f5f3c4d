#!/bin/bash
set -u
for skip in 0 1 2 3; do
echo "== skip=$skip"; EQF_TF32X3_DBG_SKIP=$skip timeout 120 python - <<'PY'
import sys, os, torch
sys.path.insert(0, '.')
from equiformer_b200 import ops
sys.path.insert(0, 'tools')
from tf32x3_microbench import timeit
dev = torch.device('cuda:0')
for (M, K, N) in [(162800, 352, 32), (97680, 384, 64), (32560, 224, 128)]:
    A = torch.randn(M, K, device=dev); Bt = torch.randn(N, K, device=dev)
    us = timeit(lambda: ops.gemm_tf32x3_raw(A, Bt))
    print(f"  {M}x{K}->{N}: {us:.1f} us, A stream {4*M*K/us/1e3:.0f} GB/s")
PY
done
