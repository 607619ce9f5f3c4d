"""Grouped small-product kernel (eqf_gemm_grouped) vs one cuBLAS call per problem, CUDA-graph timed (warm caches, no host
gaps): forward (3 problems), first-order backward (3 data + 3 weight gradients in one launch) of the node-level linears.
usage: python tools/grouped_microbench.py [atoms]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from equiformer_b200 import ops  # noqa: E402
from tools.tf32x3_microbench import timeit  # noqa: E402

LAYERS = {"merge (128,64,32 -> 128,64,32)": [(1, 128, 128), (3, 64, 64), (5, 32, 32)],
          "ffn up (128,64,32 -> 384,192,96)": [(1, 128, 384), (3, 64, 192), (5, 32, 96)],
          "ffn down (384,192,96 -> 128,64,32)": [(1, 384, 128), (3, 192, 64), (5, 96, 32)]}


def main():
    R = int(sys.argv[1]) if len(sys.argv) > 1 else 2324
    dev = torch.device("cuda:0")
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator(device=dev).manual_seed(0)
    for name, paths in LAYERS.items():
        xs = [torch.randn(R * d, ki, device=dev, generator=g) for d, ki, _ in paths]
        Ws = [torch.randn(ki, ko, device=dev, generator=g) for _, ki, ko in paths]
        gs = [torch.randn(R * d, ko, device=dev, generator=g) for d, _, ko in paths]
        outs = [torch.empty(R * d, ko, device=dev) for d, _, ko in paths]
        dxs = [torch.empty_like(x) for x in xs]
        gws = [torch.zeros_like(W) for W in Ws]
        fwd = [(0, x, W, o, 1.0, False) for x, W, o in zip(xs, Ws, outs)]
        bwd = [(1, gg, W, dx, 1.0, False) for gg, W, dx in zip(gs, Ws, dxs)] + [(2, x, gg, gw, 1.0, True) for x, gg, gw in zip(xs, gs, gws)]
        flops = sum(2 * x.shape[0] * W.shape[0] * W.shape[1] for x, W in zip(xs, Ws))
        row = {"layer": name, "atoms": R, "mflop_fwd": round(flops / 1e6, 1)}
        row["grouped_fwd_us"] = round(timeit(lambda: ops.grouped_gemm_raw(fwd)), 1)
        row["cublas_fwd_us"] = round(timeit(lambda: [x @ W for x, W in zip(xs, Ws)]), 1)
        row["grouped_bwd_us"] = round(timeit(lambda: ops.grouped_gemm_raw(bwd)), 1)
        row["grouped_dgrad_only_us"] = round(timeit(lambda: ops.grouped_gemm_raw(bwd[:3])), 1)
        row["grouped_wgrad_only_us"] = round(timeit(lambda: ops.grouped_gemm_raw(bwd[3:])), 1)
        row["cublas_bwd_us"] = round(timeit(lambda: [gg @ W.t() for gg, W in zip(gs, Ws)] + [x.t() @ gg for x, gg in zip(xs, gs)]), 1)
        row["grouped_fwd_tflops"] = round(flops / row["grouped_fwd_us"] / 1e6, 1)
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
