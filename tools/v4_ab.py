"""A/B of the round-2 'v4' switches inside ONE process on ONE GPU (box-to-box noise is larger than the effects):
EQF_FUSED_DBG_SKIP bits 8 (try_wait with a suspend-time hint), 16 (two-instruction tf32 rounding), 32 (dense vector-load k-tile
producer), 64 (entry-per-thread table build), 128 (gathers issued before the handshake; 7 / 15 = skeleton without / with the hint) for the fused DTP -> linear kernel; EQF_TF32X3_DBG_SKIP bits 8 / 16 for the
stand-alone GEMM and weight-gradient kernels.  Results must not change: every variant is compared with variant 0.
usage: python tools/v4_ab.py [fused|gemm|all] [E]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from equiformer_b200 import ops  # noqa: E402
from equiformer_b200.nets.graph_attention_transformer import DepthwiseTensorProduct  # noqa: E402
from tools.tf32x3_microbench import timeit as graph_timeit  # noqa: E402
from tools.fused_microbench import CONFIGS  # noqa: E402


def fused(E):
    dev = torch.device("cuda:0")
    for name in ("qm9_l2", "oc20_l1"):
        irreps, sh, widths = CONFIGS[name]
        plan = DepthwiseTensorProduct(irreps, sh, irreps, internal_weights=False, bias=False).tp.plan
        g = torch.Generator(device=dev).manual_seed(0)
        n_nodes = max(E // 14, 2)
        As = [torch.randn(n_nodes, 2 * l + 1, m, device=dev, generator=g) for l, m in plan.in1_blocks]
        Bs = [torch.randn(n_nodes, 2 * l + 1, m, device=dev, generator=g) for l, m in plan.in1_blocks]
        y = torch.randn(E, plan.d_y, device=dev, generator=g)
        w = torch.randn(E, plan.weight_numel, device=dev, generator=g)
        off = torch.randn(plan.weight_numel, device=dev, generator=g)
        dst = torch.sort(torch.randint(0, n_nodes, (E,), device=dev, generator=g)).values
        src = torch.randint(0, n_nodes, (E,), device=dev, generator=g)
        gat = (src, dst, Bs)
        for gi, (l, _p, K) in enumerate(plan.out_groups):
            N = widths[l][0]
            if N > ops._FUSED_MAX_N:
                continue
            Wt = torch.randn(K, N, device=dev, generator=g) / K ** 0.5
            os.environ["EQF_FUSED_DBG_SKIP"] = "0"
            ref = ops.dtp_linear_fwd_raw(plan, gi, As, y, w, Wt, gather=gat, w_offset=off).clone()
            f_un = ops.dtp_forward_raw(plan, As, y, w, gather=gat, w_offset=off)[gi]
            a2 = f_un.reshape(-1, K)
            un = ops.gemm_tf32x3_raw(a2, Wt, b_is_kn=True)
            us_gemm = graph_timeit(lambda: ops.gemm_tf32x3_raw(a2, Wt, b_is_kn=True))
            us_dtp = graph_timeit(lambda: ops.dtp_forward_raw(plan, As, y, w, gather=gat, w_offset=off))
            row = {"kernel": "dtp_linear_fwd", "config": name, "l": l, "E": E, "K": K, "N": N, "unfused_gemm_us": round(us_gemm, 1),
                   "unfused_dtp_all_groups_us": round(us_dtp, 1), "rel_diff_vs_unfused": float((ref.reshape(un.shape) - un).abs().max() / un.abs().max()),
                   "us": {}, "max_diff_vs_0": {}}
            for flags in (0, 8, 16, 1, 2, 4, 7, 0):
                os.environ["EQF_FUSED_DBG_SKIP"] = str(flags)
                out = ops.dtp_linear_fwd_raw(plan, gi, As, y, w, Wt, gather=gat, w_offset=off)
                if not (flags & 7):
                    row["max_diff_vs_0"][str(flags)] = float((out - ref).abs().max() / ref.abs().max())
                us = graph_timeit(lambda: ops.dtp_linear_fwd_raw(plan, gi, As, y, w, Wt, gather=gat, w_offset=off))
                key = str(flags) if str(flags) not in row["us"] else str(flags) + "_again"
                row["us"][key] = round(us, 1)
            os.environ["EQF_FUSED_DBG_SKIP"] = "0"
            print(json.dumps(row), flush=True)


def gemm(E):
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    shapes = [("val1_l0 (SS, N=224)", E, 224, 224), ("alpha (TS, N=128)", E, 224, 128), ("val1_l1 (TS stack 64)", 3 * E, 384, 64),
              ("val1_l2 (TS stack 32)", 5 * E, 352, 32), ("dgrad_l2 (SS, N=352)", 5 * E, 32, 352), ("rad_last (SS 960)", E, 64, 960)]
    for name, M, K, N in shapes:
        A = torch.randn(M, K, device=dev, generator=g)
        Bt = torch.randn(N, K, device=dev, generator=g)
        os.environ["EQF_TF32X3_DBG_SKIP"] = "0"
        ref = ops.gemm_tf32x3_raw(A, Bt).clone()
        row = {"kernel": "gemm_tf32x3", "shape": name, "M": M, "K": K, "N": N, "us": {}, "max_diff_vs_0": {}}
        for flags in (0, 8, 16, 24, 0):
            os.environ["EQF_TF32X3_DBG_SKIP"] = str(flags)
            out = ops.gemm_tf32x3_raw(A, Bt)
            row["max_diff_vs_0"][str(flags)] = float((out - ref).abs().max())
            us = graph_timeit(lambda: ops.gemm_tf32x3_raw(A, Bt))
            key = str(flags) if str(flags) not in row["us"] else str(flags) + "_again"
            row["us"][key] = round(us, 1)
        row["gb_s_best"] = round(4.0 * (M * K + M * N + 2 * N * K) / min(row["us"].values()) / 1e3, 1)
        os.environ["EQF_TF32X3_DBG_SKIP"] = "0"
        print(json.dumps(row), flush=True)
    for name, R, K1, N in [("wgrad l2 (TS 32)", 5 * E, 352, 32), ("wgrad l1 (TS 64)", 3 * E, 384, 64), ("wgrad alpha (SS 128)", E, 224, 128),
                           ("wgrad l0 (SS 224)", E, 224, 224)]:
        A = torch.randn(R, K1, device=dev, generator=g)
        G = torch.randn(R, N, device=dev, generator=g)
        os.environ["EQF_TF32X3_DBG_SKIP"] = "0"
        ref = ops.gemm_tf32x3_wgrad_raw(A, G).clone()
        row = {"kernel": "gemm_tf32x3_wgrad", "shape": name, "R": R, "K1": K1, "N": N, "us": {}, "max_diff_vs_0": {}}
        for flags in (0, 8, 16, 24, 0):
            os.environ["EQF_TF32X3_DBG_SKIP"] = str(flags)
            out = ops.gemm_tf32x3_wgrad_raw(A, G)
            row["max_diff_vs_0"][str(flags)] = float((out - ref).abs().max() / ref.abs().max())
            us = graph_timeit(lambda: ops.gemm_tf32x3_wgrad_raw(A, G))
            key = str(flags) if str(flags) not in row["us"] else str(flags) + "_again"
            row["us"][key] = round(us, 1)
        os.environ["EQF_TF32X3_DBG_SKIP"] = "0"
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    E = int(sys.argv[2]) if len(sys.argv) > 2 else 32560
    torch.backends.cuda.matmul.allow_tf32 = False
    if what in ("fused", "all"):
        fused(E)
    if what in ("gemm", "all"):
        gemm(E)
