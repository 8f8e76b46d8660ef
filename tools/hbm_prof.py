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
# round 2: K2 (fused SigLIP attention) at the training shape (64 images), the image kernels, the batch-1 finaliser
Ni, nh, Np, hdv = 64, 16, 256, 72
qkv_v = (torch.randn(Ni * Np, 3 * nh * hdv, device=dev) * 0.3).bfloat16()
Ov = torch.empty(Ni * Np, nh * hdv, device=dev, dtype=torch.bfloat16); Pv = torch.empty(Ni, nh, Np, Np, device=dev, dtype=torch.bfloat16)
img8 = torch.randint(0, 256, (32, 224, 224, 3), device=dev, dtype=torch.uint8); imgf = torch.empty(32, 224, 224, 3, device=dev)
augp = torch.zeros(32, 8, device=dev); augp[:, 2] = 3.0; augp[:, 3] = 0.1
acc3 = torch.randn(3, 692, 2048, device=dev); r692 = torch.randn(692, 2048, device=dev).bfloat16(); x692 = torch.empty_like(r692); y692 = torch.empty_like(r692); rs692 = torch.empty(692, device=dev)
for _ in range(2):
    ops.vit_attn_fwd(qkv_v, Ov, Pv, Ni, nh, Np, hdv, 0)
    ops.vit_attn_fwd(qkv_v, Ov, None, Ni, nh, Np, hdv, 0)
    ops.image_augment(img8, imgf, 32, 224, 224, augp)
    ops.resid_norm_fwd(r692, acc3, 3, 692 * 2048, None, x692, False, scale, None, y692, None, rs692, 692, 2048)
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
