"""Where the CTA-pair 3xTF32 kernel spends its time: skip-mode timings (EQF_TF32X3_DBG_SKIP: 1 = no transform math,
2 = no MMAs, 3 = neither) for the single-CTA and the pair kernel, then CTA 0's clock64 timeline of both."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from equiformer_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")


def timed(A, W, n=20):
    for _ in range(3):
        ops.gemm_tf32x3_raw(A, W)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0.record()
    for _ in range(n):
        ops.gemm_tf32x3_raw(A, W)
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) * 1e3 / n


for M, K, N in ((162800, 352, 32), (97680, 384, 64)):
    A = torch.randn(M, K, device=dev)
    W = torch.randn(N, K, device=dev)
    for pair in ("0", "1"):
        os.environ["EQF_TF32X3_2SM"] = pair
        row = []
        for skip in ("0", "1", "2", "3"):
            os.environ["EQF_TF32X3_DBG_SKIP"] = skip
            row.append(round(timed(A, W), 1))
        os.environ["EQF_TF32X3_DBG_SKIP"] = "0"
        print(f"[{M}x{K}->{N}] pair={pair} us: full {row[0]}  no-transform {row[1]}  no-mma {row[2]}  skeleton {row[3]}", flush=True)
    for pair in ("0", "1"):
        os.environ["EQF_TF32X3_2SM"] = pair
        dbg = torch.zeros(4 * 1024, dtype=torch.int64, device=dev)
        _lib.load().eqf_gemm_tf32x3_set_timeline(dbg.data_ptr())
        ops.gemm_tf32x3_raw(A, W)
        torch.cuda.synchronize()
        _lib.load().eqf_gemm_tf32x3_set_timeline(None)
        d = dbg.cpu().view(4, 1024)
        t0 = int(d[d > 0].min())
        names = ["producer(after empty wait)", "mma(full, a_ready, committed)", "transform(full seen, arrived)",
                 "epilogue(tmem_full seen, done)"]
        print(f"-- timeline pair={pair}")
        for r in range(4):
            v = [int(x) - t0 for x in d[r] if x > 0]
            print(" ", names[r], len(v), v[:40], "... last", v[-3:])
