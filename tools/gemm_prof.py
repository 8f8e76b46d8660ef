"""Tiny driver for ncu: runs the two dominant GEMM shapes of the LAP-3B step a few times."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lap_b200 import ops
dev = "cuda"
M, F, K = 22144, 16384, 2048
X = torch.randn(M, K, device=dev).bfloat16(); W = (torch.randn(2 * F, K, device=dev) * 0.02).bfloat16()
act = torch.empty(M, F, device=dev, dtype=torch.bfloat16); gu = torch.empty(M, 2 * F, device=dev, dtype=torch.bfloat16)
Wd = (torch.randn(K, F, device=dev) * 0.02).bfloat16(); Y = torch.empty(M, K, device=dev, dtype=torch.bfloat16)
for _ in range(3):
    ops.gemm(X, W, act, M=M, N=F, K=K, epi=ops.EPI_GEGLU, C2=gu, ldc2=2 * F)     # gate/up (dual)
    ops.gemm(act, Wd, Y, M=M, N=K, K=F, epi=ops.EPI_RESID, resid=X)              # down + residual
torch.cuda.synchronize()
print("done")
