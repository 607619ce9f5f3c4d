import sys, os
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from equiformer_b200 import _lib
dev = torch.device("cuda:0")
lib = _lib.load()
torch.set_printoptions(linewidth=250, precision=2, sci_mode=False)
for (R, K1, N, mode) in [(64, 32, 32, "ones"), (64, 32, 32, "rowid"), (64, 32, 32, "colid"), (32, 128, 64, "rand")]:
    if mode == "ones":
        A = torch.ones(R, K1, device=dev); G = torch.ones(R, N, device=dev)
    elif mode == "rowid":     # A[r, m] = 1 if r == m (identity part) -> W[m, n] = G[m, n]
        A = torch.zeros(R, K1, device=dev); A[torch.arange(K1), torch.arange(K1)] = 1.0
        G = (torch.arange(R, device=dev).float()[:, None] * 100 + torch.arange(N, device=dev).float()[None, :])
    elif mode == "colid":
        A = torch.arange(K1, device=dev).float()[None, :].repeat(R, 1) + 1; G = torch.zeros(R, N, device=dev); G[0, :] = 1.0
    else:
        A = torch.randn(R, K1, device=dev); G = torch.randn(R, N, device=dev)
    S = int(lib.eqf_gemm_tf32x3_wgrad_slices(R, K1, N))
    part = torch.full((S, K1, N), 7.0, device=dev)
    rc = lib.eqf_gemm_tf32x3_wgrad(A.data_ptr(), G.data_ptr(), part.data_ptr(), R, K1, N, K1, N, None)
    torch.cuda.synchronize()
    ref = (A.double().t() @ G.double()).float()
    out = part.sum(0)
    print(mode, "rc", rc, "slices", S, "untouched(7.0) frac", (part == 7.0).float().mean().item(), "maxdiff", (out - ref).abs().max().item())
    print(" out[:6,:10]\n", out[:6, :10].cpu()); print(" ref[:6,:10]\n", ref[:6, :10].cpu())
