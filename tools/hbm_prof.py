"""Standalone driver for ncu: the HBM-bound kernels + the fused attention kernel at LAP-3B B=32 shapes."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lap_b200 import ops
dev = "cuda"
M, D, F = 22144, 2048, 16384
x = torch.randn(M, D, device=dev).bfloat16(); y = torch.empty_like(x); rstd = torch.empty(M, device=dev)
scale = torch.randn(D, device=dev) * 0.1; dsc = torch.zeros(D, device=dev)
dy = torch.randn(M, D, device=dev).bfloat16(); dres = torch.randn(M, D, device=dev).bfloat16(); dx = torch.empty_like(x)
dact = torch.randn(M, F, device=dev).bfloat16(); gu = torch.randn(M, 2 * F, device=dev).bfloat16()
n = 400_000_000
p = torch.randn(n, device=dev); g = torch.randn(n, device=dev) * 1e-3; m = torch.zeros(n, device=dev); v = torch.zeros(n, device=dev); ema = p.clone(); w16 = torch.empty(n, device=dev, dtype=torch.bfloat16)
npart = ops.opt_num_partials(); part = torch.zeros(npart, device=dev); stats = torch.zeros(4, device=dev)
Ms, W = 16384, 1152
xs = torch.randn(Ms, W, device=dev).bfloat16(); ys = torch.empty_like(xs); mean = torch.empty(Ms, device=dev); rs = torch.empty(Ms, device=dev)
lsc, lbi = torch.ones(W, device=dev), torch.zeros(W, device=dev); dls, dlb = torch.zeros(W, device=dev), torch.zeros(W, device=dev)
B, T, NH, HD, Tpad = 32, 702, 8, 256, 704
Q = (torch.randn(B, T, NH, HD, device=dev) * 0.2).bfloat16(); Kc = torch.randn(B, Tpad, HD, device=dev).bfloat16(); Vc = torch.randn(B, Tpad, HD, device=dev).bfloat16()
bits = torch.full((B, T, Tpad // 32), -1, dtype=torch.int32, device=dev)
P = torch.empty(B, T * NH, Tpad, device=dev, dtype=torch.bfloat16); O0 = torch.empty(B * 692 * NH, HD, device=dev, dtype=torch.bfloat16); O1 = torch.empty(B * 10 * NH, HD, device=dev, dtype=torch.bfloat16)
dP = torch.randn(B, T * NH, Tpad, device=dev).bfloat16()
for _ in range(2):
    ops.rmsnorm_fwd(x, y, rstd, M, D, scale=scale)
    ops.rmsnorm_bwd(dy, x, scale, rstd, dres, dx, dsc, M, D)
    ops.geglu_bwd(dact, gu, M, F)
    ops.sumsq_partials(g, n, part)
    ops.adamw_ema(p, g, m, v, ema, w16, n, part, npart, stats, 0, n, lr=1e-4, b1=0.9, b2=0.95, eps=1e-8, wd=1e-4, bc1=0.1, bc2=0.05, clip=1.0, ema_decay=0.999, ema_on=True)
    ops.layernorm_fwd(xs, lsc, lbi, ys, mean, rs, Ms, W)
    ops.layernorm_bwd(ys, xs, lsc, mean, rs, xs, ys, dls, dlb, Ms, W)
    ops.fa_gemma_fwd(Q, Kc, Vc, bits, P, O0, O1, B, T * NH, NH, T, T, Tpad, Tpad // 32, 692 * NH, HD)
    ops.softmax_bwd(P, dP, dP, B * T * NH, Tpad)
torch.cuda.synchronize()
print("done")
