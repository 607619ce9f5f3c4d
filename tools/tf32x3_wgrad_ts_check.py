"""Parity and timing of the tensor-memory-A weight-gradient kernel (EQF_TF32X3_WGRAD_TS) against the shared-memory one.

Each variant runs in its own subprocess under a timeout (a mis-synchronised kernel hangs rather than fails).
Usage: python tools/tf32x3_wgrad_ts_check.py [out.jsonl]
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = [  # R, K1, N
    (100, 32, 32), (1000, 64, 48), (3001, 100, 64), (17, 260, 40), (11620, 32, 32), (6972, 64, 64), (2324, 128, 32),
    (162800, 352, 32), (97680, 384, 64), (32560, 224, 64), (162800, 96, 32), (97680, 192, 64),
    # 65 ... 128 columns: only differ between the variants with EQF_WGRAD_TS_LEVEL=2 (see main)
    (1000, 64, 72), (2324, 128, 128), (32560, 224, 128), (32560, 672, 128), (36000, 300, 96),
]
LEVEL = os.environ.get("EQF_WGRAD_TS_LEVEL", "1")      # "2": also route 65 ... 128 columns through tensor memory


def run_case(R, K1, N, ts):
    import torch
    from equiformer_b200 import ops
    os.environ["EQF_TF32X3_WGRAD_TS"] = LEVEL if ts else "0"
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(R + 7 * K1 + N)
    A = torch.randn(R, K1, generator=g).to(dev)
    G = torch.randn(R, N, generator=g).to(dev)
    ref = A.double().t() @ G.double()
    W = ops.gemm_tf32x3_wgrad_raw(A, G)
    torch.cuda.synchronize()
    err = float((W.double() - ref).abs().max() / ref.abs().max())
    exact = None
    if R <= 4000:
        Ai = torch.randint(-8, 9, (R, K1), generator=g).float().to(dev)
        Gi = torch.randint(-8, 9, (R, N), generator=g).float().to(dev)
        exact = bool(torch.equal(ops.gemm_tf32x3_wgrad_raw(Ai, Gi).double(), Ai.double().t() @ Gi.double()))
    for _ in range(5):
        ops.gemm_tf32x3_wgrad_raw(A, G)
    n = 30
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0.record()
    for _ in range(n):
        ops.gemm_tf32x3_wgrad_raw(A, G)
    t1.record()
    torch.cuda.synchronize()
    print(json.dumps({"R": R, "K1": K1, "N": N, "ts": ts, "rel_err": err, "exact_on_integers": exact,
                      "us": round(t0.elapsed_time(t1) * 1e3 / n, 2)}), flush=True)


def child(ts):
    sys.path.insert(0, ROOT)
    for R, K1, N in CASES:
        run_case(R, K1, N, ts)


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/tf32x3_wgrad_ts_check.jsonl"
    os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
    rc = 0
    with open(out, "w") as f:
        for ts in (0, 1):
            proc = subprocess.Popen([sys.executable, os.path.abspath(__file__), "--child", str(ts)], stdout=f,
                                    stderr=subprocess.PIPE, text=True)
            try:
                _, err = proc.communicate(timeout=150)
                if proc.returncode != 0:
                    f.write(json.dumps({"ts": ts, "error": err[-600:]}) + "\n")
                    rc = 1
            except subprocess.TimeoutExpired:
                proc.kill()
                proc.communicate()
                f.write(json.dumps({"ts": ts, "error": "timeout: the kernel hung"}) + "\n")
                rc = 1
            f.flush()
    print(open(out).read())
    return rc


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        child(int(sys.argv[2]))
    else:
        sys.exit(main())
