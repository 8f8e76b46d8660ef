import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lap_b200 import ops
dev = "cuda"
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
def bench(name, M, N, K, **kw):
    A = torch.randn(M, K, device=dev).bfloat16(); B = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    C = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for _ in range(3): ops.gemm(A, B, C, M=M, N=N, K=K, **kw)
    e0.record()
    for _ in range(10): ops.gemm(A, B, C, M=M, N=N, K=K, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{name:28s} M={M} N={N} K={K}: {ms:.3f} ms = {2*M*N*K/ms/1e9:.0f} TF", flush=True)
for (M, N, K) in [(16384, 1152, 1152), (16384, 4304, 1152), (16384, 3456, 1152), (22144, 2048, 2048)]:
    bias = torch.randn(N, device=dev); R = torch.randn(M, N, device=dev).bfloat16(); C2 = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    bench("plain", M, N, K)
    bench("plain bn128", M, N, K, block_n=128)
    bench("bias", M, N, K, bias=bias)
    bench("resid", M, N, K, epi=ops.EPI_RESID, resid=R)
    bench("bias+resid", M, N, K, epi=ops.EPI_RESID, resid=R, bias=bias)
    bench("bias_gelu(+C2)", M, N, K, epi=ops.EPI_BIAS_GELU, bias=bias, C2=C2, ldc2=N)
    bench("bias_gelu(no C2)", M, N, K, epi=ops.EPI_BIAS_GELU, bias=bias)
    bench("qscale+bias", M, N, K, epi=ops.EPI_QSCALE, bias=bias, q_cols=N // 3, q_div=8.5)
