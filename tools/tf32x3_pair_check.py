"""Parity and timing of the CTA-pair (cta_group::2) narrow-output 3xTF32 kernel against the single-CTA one.

Each kernel variant runs in its own subprocess under a timeout: a mis-synchronised pair kernel hangs rather than fails.
Usage: python tools/tf32x3_pair_check.py [out.jsonl]
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = [  # M, K, N, b_is_kn
    (256, 64, 32, 0), (128, 32, 32, 0), (100, 96, 64, 0), (130, 352, 32, 1), (1000, 100, 64, 1),
    (162800, 352, 32, 1), (97680, 384, 64, 1), (162800, 352, 32, 0), (97680, 384, 64, 0), (32560, 224, 64, 1),
]


def run_case(M, K, N, kn, pair):
    import torch
    from equiformer_b200 import ops
    os.environ["EQF_TF32X3_2SM"] = "1" if pair else "0"
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(M + 7 * K + N)
    A = torch.randn(M, K, generator=g).to(dev)
    W = (torch.randn(K, N, generator=g) if kn else torch.randn(N, K, generator=g)).to(dev)
    ref = A.double() @ (W.double() if kn else W.double().t())
    C = ops.gemm_tf32x3_raw(A, W, bool(kn))
    torch.cuda.synchronize()
    err = float((C.double() - ref).abs().max() / ref.abs().max())
    for _ in range(5):
        ops.gemm_tf32x3_raw(A, W, bool(kn))
    n = 30
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0.record()
    for _ in range(n):
        ops.gemm_tf32x3_raw(A, W, bool(kn))
    t1.record()
    torch.cuda.synchronize()
    print(json.dumps({"M": M, "K": K, "N": N, "b_is_kn": kn, "pair": pair, "rel_err": err,
                      "us": round(t0.elapsed_time(t1) * 1e3 / n, 2)}), flush=True)


def child(pair):
    sys.path.insert(0, ROOT)
    for M, K, N, kn in CASES:
        run_case(M, K, N, kn, pair)


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/tf32x3_pair_check.jsonl"
    os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
    rc = 0
    with open(out, "w") as f:
        for pair in (0, 1):
            proc = subprocess.Popen([sys.executable, os.path.abspath(__file__), "--child", str(pair)], stdout=f,
                                    stderr=subprocess.PIPE, text=True)
            try:
                _, err = proc.communicate(timeout=150)
                if proc.returncode != 0:
                    f.write(json.dumps({"pair": pair, "error": err[-600:]}) + "\n")
                    rc = 1
            except subprocess.TimeoutExpired:
                proc.kill()
                proc.communicate()
                f.write(json.dumps({"pair": pair, "error": "timeout: the kernel hung"}) + "\n")
                rc = 1
            f.flush()
    print(open(out).read())
    return rc


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        child(int(sys.argv[2]))
    else:
        sys.exit(main())
