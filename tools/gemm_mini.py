import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lap_b200 import ops
print("start", flush=True)
ops.check_device()
dev = "cuda"
for (M, N, K, cg) in [(256, 256, 512, 2), (128, 256, 64, 1), (128, 128, 64, 1), (128, 128, 64, 2), (320, 1024, 1024, 1), (200, 72, 256, 1)]:
    A = torch.randn(M, K, device=dev).bfloat16(); B = torch.randn(N, K, device=dev).bfloat16()
    C = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
    print("launch", M, N, K, cg, flush=True)
    ops.gemm(A, B, C, M=M, N=N, K=K, cta_group=cg)
    torch.cuda.synchronize()
    ref = A.float() @ B.float().T
    print("  rel", ((C.float() - ref).norm() / ref.norm()).item(), flush=True)
print("DONE")
