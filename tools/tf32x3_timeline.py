"""Timeline of CTA 0 of the tcgen05 3xTF32 GEMM (clock64 stamps per warp role) for one layer shape."""
import sys, os
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from equiformer_b200 import ops, _lib

M, K, N = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (32560, 224, 224)
dev = torch.device("cuda:0")
A = torch.randn(M, K, device=dev); Bt = torch.randn(N, K, device=dev)
for _ in range(3):
    ops.gemm_tf32x3_raw(A, Bt)
dbg = torch.zeros(4 * 1024, dtype=torch.int64, device=dev)
_lib.load().eqf_gemm_tf32x3_set_timeline(dbg.data_ptr())
ops.gemm_tf32x3_raw(A, Bt)
torch.cuda.synchronize()
_lib.load().eqf_gemm_tf32x3_set_timeline(None)
d = dbg.cpu().view(4, 1024)
t0 = int(d[d > 0].min())
names = ["producer(after empty wait)", "mma(full, lo_ready, committed)", "transform(full seen, arrived)", "epilogue(tmem_full seen, done)"]
for r in range(4):
    v = [int(x) - t0 for x in d[r] if x > 0]
    print(names[r], len(v)); print(" ", v[:48])
