"""GPU diagnostic for the tcgen05 GEMM: runs every layout/epilogue variant and prints error stats."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lap_b200 import ops

torch.manual_seed(0)
dev = "cuda"
ops.check_device()

def report(name, got, ref):
    got = got.float(); ref = ref.float()
    err = (got - ref).abs()
    rel = err.norm() / ref.norm().clamp_min(1e-30)
    bad = (err > 0.05 * ref.abs().max()).sum().item()
    print(f"{name:55s} rel={rel.item():.3e} max={err.max().item():.3e} bad={bad}/{err.numel()} nan={torch.isnan(got).sum().item()}", flush=True)
    if bad and got.dim() == 2:
        idx = (err > 0.05 * ref.abs().max()).nonzero()
        rows = idx[:, 0].unique()[:12].tolist(); cols = idx[:, 1].unique()[:12].tolist()
        print("    bad rows", rows, "cols", cols)
    return rel.item()

def run(M, N, K, a_major=0, b_major=0, **kw):
    A = torch.randn(M, K, device=dev).bfloat16()
    B = torch.randn(N, K, device=dev).bfloat16()
    ref = A.float() @ B.float().T
    Ain = A if a_major == 0 else A.T.contiguous()
    Bin = B if b_major == 0 else B.T.contiguous()
    C = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
    ops.gemm(Ain, Bin, C, M=M, N=N, K=K, a_major=a_major, b_major=b_major, **kw)
    torch.cuda.synchronize()
    return report(f"gemm M={M} N={N} K={K} maj=({a_major},{b_major}) {kw}", C, ref)

import os
CGS = [int(x) for x in os.environ.get("CGS", "2").split(",")]
for (M, N, K) in [(128, 128, 64), (128, 256, 64), (128, 128, 256), (256, 256, 512), (1024, 2048, 2048), (320, 1024, 1024), (200, 72, 256), (384, 1152, 4304), (500, 4304, 1152)]:
    for maj in [(0, 0), (0, 1), (1, 1), (1, 0)]:
        if maj[0] == 1 and M % 8: continue
        for cg in CGS:
            try:
                run(M, N, K, *maj, cta_group=cg)
            except Exception as e:
                print("EXC", M, N, K, maj, e)
run(512, 512, 512, block_n=128)
run(512, 512, 512, block_n=256)

# fp32 out + accumulate
M, N, K = 384, 512, 1000 if False else 1024
A = torch.randn(M, K, device=dev).bfloat16(); B = torch.randn(N, K, device=dev).bfloat16()
C = torch.ones(M, N, device=dev)
ops.gemm(A.T.contiguous(), B.T.contiguous(), C, M=M, N=N, K=K, a_major=1, b_major=1, accumulate=True)
report("fp32 accumulate (MN,MN)", C, A.float() @ B.float().T + 1)

# dual GEGLU
M, F, K = 300, 512, 256
X = torch.randn(M, K, device=dev).bfloat16(); W = (torch.randn(2 * F, K, device=dev) * 0.1).bfloat16()
act = torch.zeros(M, F, device=dev, dtype=torch.bfloat16); gu = torch.zeros(M, 2 * F, device=dev, dtype=torch.bfloat16)
ops.gemm(X, W, act, M=M, N=F, K=K, epi=ops.EPI_GEGLU, C2=gu, ldc2=2 * F)
g = (X.float() @ W[:F].float().T).bfloat16().float(); u = (X.float() @ W[F:].float().T).bfloat16().float()
ref = torch.nn.functional.gelu(g, approximate="tanh").bfloat16().float() * u
report("geglu act", act, ref); report("geglu g", gu[:, :F], g); report("geglu u", gu[:, F:], u)

# bias + resid + gated
M, N, K = 320, 1024, 512
A = torch.randn(M, K, device=dev).bfloat16(); B = (torch.randn(N, K, device=dev) * 0.1).bfloat16()
bias = torch.randn(N, device=dev); R = torch.randn(M, N, device=dev).bfloat16(); G = torch.randn(32, N, device=dev).bfloat16()
y = (A.float() @ B.float().T)
C = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
ops.gemm(A, B, C, M=M, N=N, K=K, bias=bias); report("bias", C, y.bfloat16().float() + bias.bfloat16().float())
ops.gemm(A, B, C, M=M, N=N, K=K, epi=ops.EPI_RESID, resid=R); report("resid", C, R.float() + y.bfloat16().float())
ops.gemm(A, B, C, M=M, N=N, K=K, epi=ops.EPI_GATED_RESID, resid=R, gate=G, ldg=N, gate_rows=10)
report("gated resid", C, R.float() + (y.bfloat16().float() * G.float().repeat_interleave(10, 0)).bfloat16().float())
C2 = torch.zeros_like(C)
ops.gemm(A, B, C, M=M, N=N, K=K, epi=ops.EPI_BIAS_GELU, bias=bias, C2=C2, ldc2=N)
pre = (y.bfloat16().float() + bias.bfloat16().float()).bfloat16().float()
report("bias_gelu pre", C2, pre); report("bias_gelu act", C, torch.nn.functional.gelu(pre, approximate="tanh"))

# batched, strided (attention-like): S[b] = Q[b] K[b]^T
Bt, T, S, H = 3, 264, 200, 256
Q = torch.randn(Bt, T, H, device=dev).bfloat16(); Kk = torch.randn(Bt, S, H, device=dev).bfloat16()
Sc = torch.zeros(Bt, T, 208, device=dev)
ops.gemm(Q, Kk, Sc, M=T, N=S, K=H, batch_i=Bt, a_bs=(T * H, 0), b_bs=(S * H, 0), c_bs=(T * 208, 0), ldc=208)
report("batched QK^T fp32", Sc[:, :, :S].reshape(-1, S), torch.einsum("bth,bsh->bts", Q.float(), Kk.float()).reshape(-1, S))
P = torch.randn(Bt, T, S, device=dev).bfloat16(); V = torch.randn(Bt, S, H, device=dev).bfloat16()
O = torch.zeros(Bt, T, H, device=dev, dtype=torch.bfloat16)
ops.gemm(P, V, O, M=T, N=H, K=S, b_major=1, batch_i=Bt, a_bs=(T * S, 0), b_bs=(S * H, 0), c_bs=(T * H, 0))
report("batched PV", O.reshape(-1, H), torch.einsum("bts,bsh->bth", P.float(), V.float()).reshape(-1, H))

# timing
for (M, N, K, kw) in [(8192, 8192, 8192, {}), (22144, 2048, 2048, {}), (22144, 2048, 16384, {}), (16384, 4304, 1152, {}), (22144, 2560, 2048, {})]:
    A = torch.randn(M, K, device=dev).bfloat16(); B = torch.randn(N, K, device=dev).bfloat16(); C = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    for cg in (1, 2):
        for _ in range(3): ops.gemm(A, B, C, M=M, N=N, K=K, cta_group=cg, **kw)
        e0.record()
        for _ in range(10): ops.gemm(A, B, C, M=M, N=N, K=K, cta_group=cg, **kw)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"   cg={cg}: {ms:.3f} ms = {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
    for _ in range(3): torch.matmul(A, B.T, out=C)
    e0.record()
    for _ in range(10): torch.matmul(A, B.T, out=C)
    e1.record(); torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / 10
    print(f"time M={M} N={N} K={K}: {ms:.3f} ms = {2*M*N*K/ms/1e9:.1f} TFLOP/s   (cublas {ms2:.3f} ms = {2*M*N*K/ms2/1e9:.1f})", flush=True)
# dual timing
M, F, K = 22144, 16384, 2048
X = torch.randn(M, K, device=dev).bfloat16(); W = (torch.randn(2 * F, K, device=dev) * 0.02).bfloat16()
act = torch.empty(M, F, device=dev, dtype=torch.bfloat16); gu = torch.empty(M, 2 * F, device=dev, dtype=torch.bfloat16)
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
for cg in (1, 2):
    for _ in range(2): ops.gemm(X, W, act, M=M, N=F, K=K, epi=ops.EPI_GEGLU, C2=gu, ldc2=2 * F, cta_group=cg)
    e0.record()
    for _ in range(5): ops.gemm(X, W, act, M=M, N=F, K=K, epi=ops.EPI_GEGLU, C2=gu, ldc2=2 * F, cta_group=cg)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"time GEGLU dual cg={cg} M={M} F={F} K={K}: {ms:.3f} ms = {2*M*2*F*K/ms/1e9:.1f} TFLOP/s")
# wgrad / dgrad shapes of the Gemma MLP
for (M, N, K, am, bm, f32) in [(22144, 16384, 2048, 0, 1, False), (32768, 2048, 22144, 1, 1, True), (2048, 16384, 22144, 1, 1, True), (22144, 2048, 32768, 0, 1, False)]:
    A = torch.randn((M, K) if am == 0 else (K, M), device=dev).bfloat16(); B = torch.randn((N, K) if bm == 0 else (K, N), device=dev).bfloat16()
    C = torch.empty(M, N, device=dev, dtype=torch.float32 if f32 else torch.bfloat16)
    for cg in (1, 2):
        for _ in range(2): ops.gemm(A, B, C, M=M, N=N, K=K, a_major=am, b_major=bm, cta_group=cg)
        e0.record()
        for _ in range(5): ops.gemm(A, B, C, M=M, N=N, K=K, a_major=am, b_major=bm, cta_group=cg)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"time maj=({am},{bm}) f32out={f32} cg={cg} M={M} N={N} K={K}: {ms:.3f} ms = {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
# split-K correctness + speed on under-filled wgrads
for (M, N, K) in [(1152, 1152, 16384), (2560, 2048, 22144), (4304, 1152, 16384), (512, 384, 4000)]:
    A = torch.randn(K, M, device=dev).bfloat16(); B = torch.randn(K, N, device=dev).bfloat16()
    ref = A.float().T @ B.float()
    for ks in (1, 0, 4):
        C = torch.full((M, N), 7.0, device=dev)
        ops.gemm(A, B, C, M=M, N=N, K=K, a_major=1, b_major=1, k_splits=ks)
        report(f"wgrad split-K ks={ks} M={M} N={N} K={K}", C, ref)
        for _ in range(2): ops.gemm(A, B, C, M=M, N=N, K=K, a_major=1, b_major=1, k_splits=ks)
        e0.record()
        for _ in range(5): ops.gemm(A, B, C, M=M, N=N, K=K, a_major=1, b_major=1, k_splits=ks)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"   ks={ks}: {ms:.3f} ms = {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
    C = torch.ones(M, N, device=dev)
    ops.gemm(A, B, C, M=M, N=N, K=K, a_major=1, b_major=1, k_splits=3, accumulate=True)
    report("wgrad split-K accumulate", C, ref + 1)
# SigLIP forward shapes with heavy epilogues
M, N, K = 16384, 4304, 1152
A = torch.randn(M, K, device=dev).bfloat16(); B = (torch.randn(N, K, device=dev) * 0.05).bfloat16(); bias = torch.randn(N, device=dev)
C = torch.empty(M, N, device=dev, dtype=torch.bfloat16); C2 = torch.empty_like(C)
for name, kw in [("bias_gelu", dict(epi=ops.EPI_BIAS_GELU, bias=bias, C2=C2, ldc2=N)), ("plain", {})]:
    for _ in range(2): ops.gemm(A, B, C, M=M, N=N, K=K, **kw)
    e0.record()
    for _ in range(5): ops.gemm(A, B, C, M=M, N=N, K=K, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"siglip fc1 {name}: {ms:.3f} ms = {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
M, N, K = 16384, 1152, 1152
A = torch.randn(M, K, device=dev).bfloat16(); B = (torch.randn(N, K, device=dev) * 0.05).bfloat16(); bias = torch.randn(N, device=dev); R = torch.randn(M, N, device=dev).bfloat16()
C = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
for _ in range(2): ops.gemm(A, B, C, M=M, N=N, K=K, bias=bias, epi=ops.EPI_RESID, resid=R)
e0.record()
for _ in range(5): ops.gemm(A, B, C, M=M, N=N, K=K, bias=bias, epi=ops.EPI_RESID, resid=R)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"siglip out-proj bias+resid: {ms:.3f} ms = {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
# leading-dimension experiment for the K=32768 dgrad shape
M, N, K = 22144, 2048, 32768
for pad in (0,):
    A = torch.randn(M, K + pad, device=dev).bfloat16(); B = torch.randn(K, N, device=dev).bfloat16(); C = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for _ in range(2): ops.gemm(A, B, C, M=M, N=N, K=K, b_major=1, lda=K + pad)
    e0.record()
    for _ in range(5): ops.gemm(A, B, C, M=M, N=N, K=K, b_major=1, lda=K + pad)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"dgrad K=32768 lda pad {pad}: {ms:.3f} ms = {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
for blk in ():
    A = torch.randn(M, K, device=dev).bfloat16(); B = torch.randn(K, N, device=dev).bfloat16(); C = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for _ in range(2): ops.gemm(A, B, C, M=M, N=N, K=K, b_major=1, block_n=blk)
    e0.record()
    for _ in range(5): ops.gemm(A, B, C, M=M, N=N, K=K, b_major=1, block_n=blk)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"dgrad K=32768 block_n {blk}: {ms:.3f} ms = {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
print("DONE")
