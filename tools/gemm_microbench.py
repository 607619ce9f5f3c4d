"""Time the fast-fp32 tcgen05 GEMM against cuBLAS SGEMM on the per-degree linear shapes of one QM9 layer."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from equiformer_b200 import ops  # noqa: E402


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) * 1e3 / iters


def main():
    E = int(sys.argv[1]) if len(sys.argv) > 1 else 32560
    dev = torch.device("cuda:0")
    torch.backends.cuda.matmul.allow_tf32 = False
    only_wgrad = os.environ.get("ONLY_WGRAD", "0") == "1"
    shapes = [("val1_l0", E, 224, 224), ("alpha", E, 224, 128), ("val1_l1", 3 * E, 384, 64), ("val1_l2", 5 * E, 352, 32),
              ("val2_l0", E, 224, 128), ("rad_last", E, 64, 960), ("rad_first", E, 128, 64), ("node_l0", 2324, 128, 128)]
    g = torch.Generator(device=dev).manual_seed(0)
    for name, M, K, N in shapes:
        A = torch.randn(M, K, device=dev, generator=g)
        B = torch.randn(K, N, device=dev, generator=g)
        dC = torch.randn(M, N, device=dev, generator=g)
        row = {"shape": name, "M": M, "K": K, "N": N}
        ref = {0: A.double() @ B.double(), 1: dC.double() @ B.double().t(), 2: A.double().t() @ dC.double()}
        for mode, (a, b) in {0: (A, B), 1: (dC, B), 2: (A, dC)}.items():
            if only_wgrad and mode != 2:
                continue
            out = ops.gemm_raw(mode, a, b)
            err = ((out.double() - ref[mode]).abs().max() / ref[mode].abs().max()).item()
            us = timeit(lambda: ops.gemm_raw(mode, a, b))
            tfn = {0: lambda: a @ b, 1: lambda: a @ b.t(), 2: lambda: a.t() @ b}[mode]
            us_t = timeit(tfn)
            err_t = ((tfn().double() - ref[mode]).abs().max() / ref[mode].abs().max()).item()
            nbytes = 4 * (a.numel() + b.numel() + out.numel())
            row[f"mode{mode}"] = {"us": round(us, 1), "cublas_us": round(us_t, 1), "gb_s": round(nbytes / us / 1e3, 1),
                                  "tflops": round(2 * M * N * K / us / 1e6, 1), "rel_err": float(f"{err:.2e}"),
                                  "cublas_err": float(f"{err_t:.2e}")}
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
