"""GPU tests of individual kernels through the C ABI, each against a plain PyTorch fp32 reference of the same op."""
import os
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from lap_b200 import ops  # noqa: E402
from tests.helpers import rel_err  # noqa: E402

DEV = "cuda"


def _gemm_ref(A, B):
    return A.float() @ B.float().T


@pytest.mark.parametrize("maj", [(0, 0), (0, 1), (1, 1), (1, 0)])
@pytest.mark.parametrize("shape", [(128, 128, 64), (320, 1024, 1024), (200, 72, 256), (384, 1152, 4304), (1000, 2560, 2048)])
def test_gemm_layouts(maj, shape):
    M, N, K = shape
    torch.manual_seed(M + N + K)
    A = torch.randn(M, K, device=DEV).bfloat16()
    B = torch.randn(N, K, device=DEV).bfloat16()
    Ain = A if maj[0] == 0 else A.T.contiguous()
    Bin = B if maj[1] == 0 else B.T.contiguous()
    C = torch.zeros(M, N, device=DEV, dtype=torch.float32)
    ops.gemm(Ain, Bin, C, M=M, N=N, K=K, a_major=maj[0], b_major=maj[1])
    # fp32 output of a bf16 x bf16 GEMM with fp32 accumulation: exact up to summation order
    assert rel_err(C, _gemm_ref(A, B)) < 1e-5


def test_gemm_bf16_output_is_correctly_rounded():
    M, N, K = 256, 512, 512
    A = torch.randn(M, K, device=DEV).bfloat16()
    B = torch.randn(N, K, device=DEV).bfloat16()
    C = torch.zeros(M, N, device=DEV, dtype=torch.bfloat16)
    ops.gemm(A, B, C, M=M, N=N, K=K)
    ref = _gemm_ref(A, B)
    ulp = (C.float() - ref).abs() / ref.abs().clamp_min(1e-3)
    assert ulp.max() < 2 ** -7  # within one bf16 ulp everywhere
    assert (C == ref.bfloat16()).float().mean() > 0.99


def test_gemm_epilogues():
    M, N, K = 320, 1024, 512
    A = torch.randn(M, K, device=DEV).bfloat16()
    B = (torch.randn(N, K, device=DEV) * 0.1).bfloat16()
    bias = torch.randn(N, device=DEV)
    R = torch.randn(M, N, device=DEV).bfloat16()
    G = torch.randn(32, N, device=DEV).bfloat16()
    y = _gemm_ref(A, B)
    yb = y.bfloat16().float()
    C = torch.zeros(M, N, device=DEV, dtype=torch.bfloat16)
    ops.gemm(A, B, C, M=M, N=N, K=K, bias=bias)
    assert rel_err(C, (yb + bias.bfloat16().float()).bfloat16()) < 1e-3
    ops.gemm(A, B, C, M=M, N=N, K=K, epi=ops.EPI_RESID, resid=R)
    assert rel_err(C, (R.float() + yb).bfloat16()) < 1e-3
    Y2 = torch.zeros_like(C)
    ops.gemm(A, B, C, M=M, N=N, K=K, epi=ops.EPI_GATED_RESID, resid=R, gate=G, ldg=N, gate_rows=10, C2=Y2, ldc2=N)
    gate = G.float().repeat_interleave(10, 0)
    assert rel_err(C, (R.float() + (yb * gate).bfloat16().float()).bfloat16()) < 1e-3
    assert rel_err(Y2, yb) < 1e-3
    C2 = torch.zeros_like(C)
    ops.gemm(A, B, C, M=M, N=N, K=K, epi=ops.EPI_BIAS_GELU, bias=bias, C2=C2, ldc2=N)
    pre = (yb + bias.bfloat16().float()).bfloat16().float()
    assert rel_err(C2, pre) < 1e-3
    assert rel_err(C, torch.nn.functional.gelu(pre, approximate="tanh")) < 3e-3
    ops.gemm(A, B, C, M=M, N=N, K=K, epi=ops.EPI_QSCALE, q_cols=512, q_div=8.5)
    exp = yb.clone()
    exp[:, :512] /= 8.5
    assert rel_err(C, exp) < 3e-3
    # fp32 accumulate into C
    Cf = torch.ones(M, N, device=DEV)
    ops.gemm(A, B, Cf, M=M, N=N, K=K, accumulate=True)
    assert rel_err(Cf, y + 1) < 1e-5


def test_gemm_geglu_dual():
    M, F, K = 300, 512, 256
    X = torch.randn(M, K, device=DEV).bfloat16()
    W = (torch.randn(2 * F, K, device=DEV) * 0.1).bfloat16()
    act = torch.zeros(M, F, device=DEV, dtype=torch.bfloat16)
    gu = torch.zeros(M, 2 * F, device=DEV, dtype=torch.bfloat16)
    ops.gemm(X, W, act, M=M, N=F, K=K, epi=ops.EPI_GEGLU, C2=gu, ldc2=2 * F)
    g = _gemm_ref(X, W[:F]).bfloat16().float()
    u = _gemm_ref(X, W[F:]).bfloat16().float()
    assert rel_err(gu[:, :F], g) < 1e-3 and rel_err(gu[:, F:], u) < 1e-3
    assert rel_err(act, torch.nn.functional.gelu(g, approximate="tanh").bfloat16().float() * u) < 3e-3


def test_gemm_batched_broadcast_and_strided():
    Bt, T, S, H = 3, 264, 200, 256
    Q = torch.randn(Bt, T, H, device=DEV).bfloat16()
    Kk = torch.randn(Bt, S, H, device=DEV).bfloat16()
    Sc = torch.zeros(Bt, T, 208, device=DEV)
    ops.gemm(Q, Kk, Sc, M=T, N=S, K=H, batch_i=Bt, a_bs=(T * H, 0), b_bs=(S * H, 0), c_bs=(T * 208, 0), ldc=208)
    assert rel_err(Sc[:, :, :S], torch.einsum("bth,bsh->bts", Q.float(), Kk.float())) < 2e-6
    # shared (broadcast) B across the batch
    W = torch.randn(128, H, device=DEV).bfloat16()
    O = torch.zeros(Bt, T, 128, device=DEV)
    ops.gemm(Q, W, O, M=T, N=128, K=H, batch_i=Bt, a_bs=(T * H, 0), b_bs=(0, 0), c_bs=(T * 128, 0))
    assert rel_err(O, Q.float() @ W.float().T) < 2e-6


def test_gemm_rejects_bad_arguments():
    A = torch.zeros(128, 64, device=DEV, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError):
        ops.gemm(A, A, torch.zeros(128, 12, device=DEV), M=128, N=12, K=64)  # N % 8
    with pytest.raises(RuntimeError):
        ops.gemm(A, A, torch.zeros(128, 128, device=DEV, dtype=torch.bfloat16), M=128, N=128, K=64, accumulate=True)


def test_layernorm_fwd_bwd():
    M, W = 300, 1152
    x = (torch.randn(M, W, device=DEV) * 2 + 0.5).bfloat16()
    sc, bi = torch.randn(W, device=DEV), torch.randn(W, device=DEV)
    y = torch.empty_like(x)
    mean, rstd = torch.empty(M, device=DEV), torch.empty(M, device=DEV)
    ops.layernorm_fwd(x, sc, bi, y, mean, rstd, M, W)
    xr = x.float().requires_grad_(True)
    scr, bir = sc.clone().requires_grad_(True), bi.clone().requires_grad_(True)
    yr = torch.nn.functional.layer_norm(xr, (W,), scr, bir, eps=1e-6)
    assert rel_err(y, yr) < 3e-3
    dy = torch.randn(M, W, device=DEV).bfloat16()
    dres = torch.randn(M, W, device=DEV).bfloat16()
    dx = torch.empty_like(x)
    dsc, dbi = torch.zeros(W, device=DEV), torch.zeros(W, device=DEV)
    ops.layernorm_bwd(dy, x, sc, mean, rstd, dres, dx, dsc, dbi, M, W)
    yr.backward(dy.float())
    assert rel_err(dx, xr.grad + dres.float()) < 4e-3
    assert rel_err(dsc, scr.grad) < 1e-3 and rel_err(dbi, bir.grad) < 1e-3


def test_rmsnorm_plain_and_adaptive():
    B, A, D = 4, 10, 1024
    M = B * A
    x = torch.randn(M, D, device=DEV).bfloat16()
    sc = torch.randn(D, device=DEV) * 0.1
    y, rstd = torch.empty_like(x), torch.empty(M, device=DEV)
    ops.rmsnorm_fwd(x, y, rstd, M, D, scale=sc)
    xr = x.float().requires_grad_(True)
    scr = sc.clone().requires_grad_(True)
    yr = xr * torch.rsqrt((xr * xr).mean(-1, keepdim=True) + 1e-6) * (1 + scr)
    assert rel_err(y, yr) < 3e-3
    dy, dres = torch.randn(M, D, device=DEV).bfloat16(), torch.randn(M, D, device=DEV).bfloat16()
    dx, dsc = torch.empty_like(x), torch.zeros(D, device=DEV)
    ops.rmsnorm_bwd(dy, x, sc, rstd, dres, dx, dsc, M, D)
    yr.backward(dy.float())
    assert rel_err(dx, xr.grad + dres.float()) < 4e-3 and rel_err(dsc, scr.grad) < 1e-3
    # adaptive
    mod = (torch.randn(B, 3 * D, device=DEV) * 0.2).bfloat16()
    ops.rmsnorm_fwd(x, y, rstd, M, D, mod=mod, ldmod=3 * D, rows_per_sample=A)
    xr = x.float().requires_grad_(True)
    mr = mod.float().requires_grad_(True)
    s_, sh_, _ = mr[:, None, :].chunk(3, -1)
    n = (xr * torch.rsqrt((xr * xr).mean(-1, keepdim=True) + 1e-6)).view(B, A, D)
    yr = (n * (1 + s_) + sh_).view(M, D)
    assert rel_err(y, yr) < 4e-3
    dmod = torch.zeros(B, 3 * D, device=DEV, dtype=torch.bfloat16)
    ops.ada_rmsnorm_bwd(dy, x, mod, 3 * D, rstd, None, dx, dmod, 3 * D, B, A, D)
    yr.backward(dy.float())
    assert rel_err(dx, xr.grad) < 6e-3
    assert rel_err(dmod[:, : 2 * D], mr.grad[:, : 2 * D]) < 6e-3


@pytest.mark.parametrize("M,D", [(4099, 2048), (2048, 1152), (1500, 72)])
def test_forward_norms_warp_per_row_path(M, D):
    """M >= 1024 contiguous rows take the warp-per-row kernels (rows kept in registers); same arithmetic as the
    CTA-per-row kernels used for small M / gathered rows."""
    x = (torch.randn(M, D, device=DEV) * 1.5 + 0.25).bfloat16()
    sc, bi = torch.randn(D, device=DEV) * 0.3, torch.randn(D, device=DEV)
    y, mean, rstd = torch.empty_like(x), torch.empty(M, device=DEV), torch.empty(M, device=DEV)
    ops.rmsnorm_fwd(x, y, rstd, M, D, scale=sc)
    xf = x.float()
    r = torch.rsqrt((xf * xf).mean(-1, keepdim=True) + 1e-6)
    assert rel_err(y, xf * r * (1 + sc)) < 3e-3 and rel_err(rstd, r[:, 0]) < 1e-5
    ops.layernorm_fwd(x, sc, bi, y, mean, rstd, M, D)
    assert rel_err(y, torch.nn.functional.layer_norm(xf, (D,), sc, bi, eps=1e-6)) < 3e-3
    assert rel_err(mean, xf.mean(-1)) < 1e-4
    # small-M path gives the same values on the same rows
    y2, rstd2 = torch.empty(512, D, device=DEV, dtype=torch.bfloat16), torch.empty(512, device=DEV)
    ops.rmsnorm_fwd(x[:512].contiguous(), y2, rstd2, 512, D, scale=sc)
    ops.rmsnorm_fwd(x, y, rstd, M, D, scale=sc)
    assert rel_err(y[:512], y2) < 1e-3 and rel_err(rstd[:512], rstd2) < 1e-6


def test_rope_roundtrip_and_reference():
    from oracle.lap_oracle import apply_rope

    B, P, A, NH, HD = 2, 20, 4, 8, 256
    T = P + A
    Tpad = 32
    ld = (NH + 2) * HD
    qkv0 = torch.randn(B * P, ld, device=DEV).bfloat16()
    qkv1 = torch.randn(B * A, ld, device=DEV).bfloat16()
    pos = torch.randint(0, 700, (B, T), device=DEV, dtype=torch.int32)
    ts = (10_000.0 ** ((2.0 / HD) * torch.arange(HD // 2, dtype=torch.float32))).to(DEV)
    Q = torch.zeros(B, T, NH, HD, device=DEV, dtype=torch.bfloat16)
    Kc = torch.zeros(B, Tpad, HD, device=DEV, dtype=torch.bfloat16)
    Vc = torch.zeros_like(Kc)
    ops.rope_fwd(qkv0, qkv1, pos, ts, Q, Kc, Vc, B, P, A, Tpad, NH, HD, 0, HD ** -0.5)
    allq = torch.cat([qkv0.view(B, P, ld), qkv1.view(B, A, ld)], 1).float().cpu()
    q_ref = apply_rope(allq[..., : NH * HD].reshape(B, T, NH, HD), pos.cpu(), True) * HD ** -0.5
    k_ref = apply_rope(allq[..., NH * HD:(NH + 1) * HD].reshape(B, T, 1, HD), pos.cpu(), True)[:, :, 0]
    assert rel_err(Q, q_ref) < 2e-3
    assert rel_err(Kc[:, :T], k_ref) < 2e-3
    assert torch.equal(Vc[:, :T].cpu().float(), allq[..., (NH + 1) * HD:])
    assert Kc[:, T:].abs().max() == 0
    # backward of a rotation is the inverse rotation: rope_bwd(rope_fwd(x)) == x * qscale^2 for q, x for k
    d0, d1 = torch.zeros_like(qkv0), torch.zeros_like(qkv1)
    ops.rope_bwd(Q, Kc, Vc, pos, ts, d0, d1, B, P, A, Tpad, NH, HD, HD ** -0.5)
    assert rel_err(d0[:, : NH * HD], qkv0[:, : NH * HD].float() / HD) < 8e-3
    assert rel_err(d0[:, NH * HD:(NH + 1) * HD], qkv0[:, NH * HD:(NH + 1) * HD]) < 8e-3


def test_attn_softmax_and_fully_masked_rows():
    B, Tq, G, T, Tpad = 2, 5, 8, 70, 96
    R = Tq * G
    S = torch.randn(B, R, Tpad, device=DEV)
    dense = torch.rand(B, Tq, T, device=DEV) < 0.6
    dense[0, 0] = False  # fully masked query row -> uniform probabilities (gemma.py:258 finite big_neg)
    W32 = Tpad // 32
    bits = torch.zeros(B, Tq, W32, dtype=torch.int64, device=DEV)
    for j in range(T):
        bits[:, :, j // 32] |= dense[:, :, j].long() << (j % 32)
    bits32 = (bits & 0xFFFFFFFF).to(torch.int64)
    bits32 = torch.where(bits32 >= 2 ** 31, bits32 - 2 ** 32, bits32).to(torch.int32)
    Pm = torch.empty(B, R, Tpad, device=DEV, dtype=torch.bfloat16)
    ops.attn_softmax_fwd(S, bits32, Pm, B, R, G, T, Tpad, W32)
    m = dense.repeat_interleave(G, 1)
    ref = torch.softmax(torch.where(m, S[:, :, :T], torch.tensor(-2.3819763e38, device=DEV)), -1)
    assert rel_err(Pm[:, :, :T], ref) < 3e-3
    assert Pm[:, :, T:].abs().max() == 0
    assert torch.allclose(Pm[0, 0, :T].float(), torch.full((T,), 1.0 / T, device=DEV), rtol=1e-2)
    dP = torch.randn(B, R, Tpad, device=DEV).bfloat16()
    dS = torch.empty_like(dP)
    ops.softmax_bwd(Pm, dP, dS, B * R, Tpad)
    p, d = Pm.float(), dP.float()
    assert rel_err(dS, p * (d - (p * d).sum(-1, keepdim=True))) < 4e-3


def test_mask_build_bit_exact_random():
    from oracle import lap_oracle as O

    rng = np.random.default_rng(0)
    B, n_img, L, A = 6, 32, 24, 10
    P_ = n_img + L
    T = P_ + A
    Tpad = (T + 31) // 32 * 32
    pm = np.concatenate([np.repeat(rng.random((B, 1)) < 0.8, n_img, 1), np.arange(L)[None] < rng.integers(0, L + 1, (B, 1))], 1)
    la_text = np.zeros((B, L), bool)
    for b in range(B):
        s = rng.integers(1, L - 2)
        la_text[b, s:s + rng.integers(0, 8)] = True
    la_text &= pm[:, n_img:]
    par = np.concatenate([np.zeros((B, n_img), bool), la_text], 1)
    pma = pm & ~par
    sm = np.ones((B, A), bool)
    sar = np.tile(np.array([1] + [0] * (A - 1), bool), (B, 1))
    up = lambda x: torch.from_numpy(x.astype(np.uint8)).to(DEV)
    bits = torch.zeros(B, T, Tpad // 32, dtype=torch.int32, device=DEV)
    pos = torch.zeros(B, T, dtype=torch.int32, device=DEV)
    ops.mask_build(up(pm), up(par), up(pma), up(sm), up(sar), bits, pos, B, P_, A, Tpad // 32)
    dense = torch.zeros(B, T, T, dtype=torch.uint8, device=DEV)
    ops.mask_expand(bits, dense, B * T, T, Tpad // 32)
    t = torch.from_numpy
    ref = O.build_combined_attention_mask(t(pm), t(par), t(pma), t(sm), t(sar))
    refpos = O.build_combined_positions(t(pm), t(pma), t(sm))
    assert torch.equal(dense.cpu().bool(), ref)
    assert torch.equal(pos.cpu(), refpos)


def test_ce_loss_and_gradient():
    R, V = 64, 4096
    logits = torch.randn(R, V, device=DEV) * 3
    tgt = torch.randint(0, V, (R,), device=DEV, dtype=torch.int32)
    w = torch.rand(R, device=DEV)
    w[-5:] = 0
    nll = torch.empty(R, device=DEV)
    dl = torch.empty(R, V, device=DEV, dtype=torch.bfloat16)
    ops.ce_fwd_bwd(logits, V, tgt, w, nll, dl, V, R, V)
    lr = logits.clone().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(lr, tgt.long(), reduction="none")
    assert rel_err(nll, ref) < 1e-5
    (ref * w).sum().backward()
    assert rel_err(dl, lr.grad) < 3e-3
    assert dl[-5:].abs().max() == 0


def test_adamw_ema_matches_closed_form():
    n = 4096 * 3
    p = torch.randn(n, device=DEV)
    g = torch.randn(n, device=DEV) * 3
    m, v = torch.rand(n, device=DEV) * 0.1, torch.rand(n, device=DEV) * 0.1
    ema = p.clone()
    p0, m0, v0 = p.clone(), m.clone(), v.clone()
    w16 = torch.empty(n, device=DEV, dtype=torch.bfloat16)
    npart = ops.opt_num_partials()
    part = torch.zeros(npart, device=DEV)
    stats = torch.zeros(4, device=DEV)
    ops.sumsq_partials(g, n, part)
    hp = dict(lr=1e-2, b1=0.9, b2=0.95, eps=1e-8, wd=1e-4, bc1=1 - 0.9 ** 3, bc2=1 - 0.95 ** 3, clip=1.0, ema_decay=0.99, ema_on=True)
    ops.adamw_ema(p, g, m, v, ema, w16, n, part, npart, stats, 0, 4096, **hp)
    gn = g.double().norm().float()
    gs = g * (1.0 / gn)
    mr = 0.9 * m0 + 0.1 * gs
    vr = 0.95 * v0 + 0.05 * gs * gs
    pr = p0 - 1e-2 * ((mr / hp["bc1"]) / ((vr / hp["bc2"]).sqrt() + 1e-8) + 1e-4 * p0)
    assert abs(stats[0].item() - gn.item()) < 1e-3 * gn.item()
    assert torch.allclose(p, pr, rtol=1e-5, atol=1e-7) and torch.allclose(m, mr, rtol=1e-5, atol=1e-8)
    assert torch.allclose(ema, 0.99 * p0 + 0.01 * pr, rtol=1e-5, atol=1e-7)
    assert torch.equal(w16, pr.bfloat16()) or rel_err(w16, pr) < 3e-3
    assert abs(stats[2].sqrt().item() - pr[:4096].norm().item()) < 1e-3 * pr[:4096].norm().item()


def test_split_hi_lo_recovers_fp32_table():
    E = torch.randn(1000, 64, device=DEV) * 0.02
    out = torch.empty(1000, 128, device=DEV, dtype=torch.bfloat16)
    ops.split_hi_lo(E, out, 1000, 64)
    rec = out[:, :64].float() + out[:, 64:].float()
    assert ((rec - E).abs() / E.abs().clamp_min(1e-6)).max() < 2 ** -15


def test_sgemm_patchify_colsum_embed():
    B, C, H, ps, W = 2, 2, 56, 14, 32
    imgs = [torch.rand(B, H, H, 3, device=DEV) * 2 - 1 for _ in range(C)]
    np_ = (H // ps) ** 2
    pk = ps * ps * 3
    patches = torch.empty(B * C * np_, pk, device=DEV)
    ops.patchify(imgs, patches, B, C, H, H, ps)
    ref = torch.stack(imgs, 1).reshape(B * C, H // ps, ps, H // ps, ps, 3).permute(0, 1, 3, 2, 4, 5).reshape(-1, pk)
    assert torch.equal(patches, ref)
    u8 = [(torch.rand(B, H, H, 3, device=DEV) * 255).to(torch.uint8) for _ in range(C)]
    ops.patchify(u8, patches, B, C, H, H, ps)
    refu = (torch.stack(u8, 1).float() / 255.0 * 2.0 - 1.0).reshape(B * C, H // ps, ps, H // ps, ps, 3).permute(0, 1, 3, 2, 4, 5).reshape(-1, pk)
    assert torch.allclose(patches, refu, atol=1e-6)
    Wk, bias, pos = torch.randn(W, pk, device=DEV), torch.randn(W, device=DEV), torch.randn(np_, W, device=DEV)
    out = torch.empty(B * C * np_, W, device=DEV)
    ops.sgemm(ref, Wk, out, B * C * np_, W, pk, pk, 1, pk, 1, ldc=W, bias=bias, table=pos, table_rows=np_)
    exp = ref @ Wk.T + bias + pos.repeat(B * C, 1)
    assert rel_err(out, exp) < 1e-5
    X = torch.randn(500, 96, device=DEV).bfloat16()
    cs = torch.zeros(96, device=DEV)
    ops.colsum(X, 96, cs, 500, 96)
    assert rel_err(cs, X.float().sum(0)) < 1e-4
    ids = torch.randint(0, 100, (2, 7), device=DEV, dtype=torch.int32)
    E = torch.randn(100, 64, device=DEV)
    Xo = torch.zeros(2 * 12, 64, device=DEV, dtype=torch.bfloat16)
    ops.embed_fwd(ids, E, Xo, 2, 7, 5, 12, 64, 8.0)
    assert rel_err(Xo.view(2, 12, 64)[:, 5:], (E[ids.long()] * 8.0).bfloat16()) == 0
    dE = torch.zeros_like(E)
    ops.embed_bwd(ids, Xo, dE, 2, 7, 5, 12, 64, 8.0)
    ref_dE = torch.zeros_like(E).index_add_(0, ids.view(-1).long(), Xo.view(2, 12, 64)[:, 5:].reshape(-1, 64).float() * 8.0)
    assert rel_err(dE, ref_dE) < 1e-5


@pytest.mark.parametrize("M,N,K", [(10, 1024, 4096), (1, 3072, 1024), (16, 160, 64), (7, 2560, 1024)])
def test_skinny_gemm_epilogues(M, N, K):
    torch.manual_seed(M * N)
    X = torch.randn(M, K, device=DEV).bfloat16()
    W = (torch.randn(N, K, device=DEV) * 0.05).bfloat16()
    ref = X.float() @ W.float().T
    Y = torch.zeros(M, N, device=DEV, dtype=torch.bfloat16)
    ops.skinny_gemm(X, W, Y, M=M, N=N, K=K)
    assert rel_err(Y, ref.bfloat16()) < 1e-3
    bias = torch.randn(N, device=DEV)
    ops.skinny_gemm(X, W, Y, M=M, N=N, K=K, bias=bias)
    assert rel_err(Y, (ref.bfloat16().float() + bias.bfloat16().float()).bfloat16()) < 1e-3
    Yf = torch.zeros(M, N, device=DEV)
    ops.skinny_gemm(X, W, Yf, M=M, N=N, K=K, bias=bias)
    assert rel_err(Yf, ref + bias) < 1e-5
    R = torch.randn(M, N, device=DEV).bfloat16()
    ops.skinny_gemm(X, W, Y, M=M, N=N, K=K, epi=ops.EPI_RESID, resid=R)
    assert rel_err(Y, (R.float() + ref.bfloat16().float()).bfloat16()) < 1e-3
    G = torch.randn(2, N, device=DEV).bfloat16()
    Y2 = torch.zeros_like(Y)
    rows = (M + 1) // 2
    ops.skinny_gemm(X, W, Y, M=M, N=N, K=K, epi=ops.EPI_GATED_RESID, resid=R, gate=G, ldg=N, gate_rows=rows, Y2=Y2, ldy2=N)
    gate = G.float().repeat_interleave(rows, 0)[:M]
    assert rel_err(Y, (R.float() + (ref.bfloat16().float() * gate).bfloat16().float()).bfloat16()) < 1e-3
    assert rel_err(Y2, ref.bfloat16()) < 1e-3
    if N % 16 == 0:
        F = N // 2
        act = torch.zeros(M, F, device=DEV, dtype=torch.bfloat16)
        gu = torch.zeros(M, N, device=DEV, dtype=torch.bfloat16)
        ops.skinny_gemm(X, W, act, M=M, N=F, K=K, epi=ops.EPI_GEGLU, Y2=gu, ldy2=N)
        g = ref[:, :F].bfloat16().float()
        u = ref[:, F:].bfloat16().float()
        assert rel_err(gu, torch.cat([g, u], 1)) < 1e-3
        assert rel_err(act, torch.nn.functional.gelu(g, approximate="tanh").bfloat16().float() * u) < 3e-3


def test_decode_attention_matches_reference():
    B, Tq, NH, HD, T, Tpad = 2, 10, 8, 256, 702, 704
    W32 = Tpad // 32
    Q = (torch.randn(B, Tq, NH, HD, device=DEV) * 0.1).bfloat16()
    Kc = torch.randn(B, Tpad, HD, device=DEV).bfloat16()
    Vc = torch.randn(B, Tpad, HD, device=DEV).bfloat16()
    dense = torch.rand(B, Tq, T, device=DEV) < 0.7
    dense[1, 3] = False
    bits = torch.zeros(B, Tq, W32, dtype=torch.int64, device=DEV)
    for j in range(T):
        bits[:, :, j // 32] |= dense[:, :, j].long() << (j % 32)
    bits32 = torch.where(bits >= 2 ** 31, bits - 2 ** 32, bits).to(torch.int32)
    O = torch.zeros(B, Tq, NH, HD, device=DEV, dtype=torch.bfloat16)
    ops.decode_attn(Q, Kc, Vc, bits32, O, B, Tq, NH, HD, T, Tpad, W32)
    logits = torch.einsum("bqhd,bsd->bhqs", Q.float(), Kc[:, :T].float())
    logits = torch.where(dense[:, None], logits, torch.tensor(-2.3819763e38, device=DEV))
    p = torch.softmax(logits, -1).bfloat16().float()
    ref = torch.einsum("bhqs,bsd->bqhd", p, Vc[:, :T].float())
    assert rel_err(O, ref) < 4e-3


@pytest.mark.parametrize("B,Tq,T,split_tok", [(2, 178, 178, 168), (1, 702, 702, 692), (2, 10, 178, 0)])
def test_fused_attention_matches_unfused(B, Tq, T, split_tok):
    """K1 fused tcgen05 attention vs GEMM + masked softmax + GEMM on the same inputs (same rounding points)."""
    NH, HD = 8, 256
    Tpad = (T + 63) // 64 * 64
    W32 = Tpad // 32
    R = Tq * NH
    torch.manual_seed(T + Tq)
    Q = (torch.randn(B, Tq, NH, HD, device=DEV) * 0.25).bfloat16()
    Kc = torch.zeros(B, Tpad, HD, device=DEV, dtype=torch.bfloat16)
    Vc = torch.zeros_like(Kc)
    Kc[:, :T] = torch.randn(B, T, HD, device=DEV).bfloat16()
    Vc[:, :T] = torch.randn(B, T, HD, device=DEV).bfloat16()
    dense = torch.rand(B, Tq, T, device=DEV) < 0.6
    dense[0, 1] = False  # fully masked row
    bits = torch.zeros(B, Tq, W32, dtype=torch.int64, device=DEV)
    for j in range(T):
        bits[:, :, j // 32] |= dense[:, :, j].long() << (j % 32)
    bits32 = torch.where(bits >= 2 ** 31, bits - 2 ** 32, bits).to(torch.int32)
    # unfused
    S = torch.zeros(B, R, Tpad, device=DEV)
    ops.gemm(Q, Kc, S, M=R, N=Tpad, K=HD, ldc=Tpad, batch_i=B, a_bs=(R * HD, 0), b_bs=(Tpad * HD, 0), c_bs=(R * Tpad, 0))
    P_ref = torch.zeros(B, R, Tpad, device=DEV, dtype=torch.bfloat16)
    ops.attn_softmax_fwd(S, bits32, P_ref, B, R, NH, T, Tpad, W32)
    O_ref = torch.zeros(B, R, HD, device=DEV, dtype=torch.bfloat16)
    ops.gemm(P_ref, Vc, O_ref, M=R, N=HD, K=Tpad, b_major=1, lda=Tpad, ldb=HD, ldc=HD, batch_i=B,
             a_bs=(R * Tpad, 0), b_bs=(Tpad * HD, 0), c_bs=(R * HD, 0))
    # fused
    split = split_tok * NH
    P = torch.full((B, R, Tpad), 7.0, device=DEV, dtype=torch.bfloat16)
    O0 = torch.zeros(B, max(split, 1), HD, device=DEV, dtype=torch.bfloat16)
    O1 = torch.zeros(B, max(R - split, 1), HD, device=DEV, dtype=torch.bfloat16)
    ops.fa_gemma_fwd(Q, Kc, Vc, bits32, P, O0, O1, B, R, NH, Tq, T, Tpad, W32, split, HD)
    torch.cuda.synchronize()
    assert rel_err(P, P_ref) < 2e-3
    assert (P.float() - P_ref.float()).abs().max() < 2 ** -7  # at most one bf16 ulp of a probability
    O = torch.cat([O0[:, :split], O1[:, : R - split]], 1) if 0 < split < R else (O0 if split == R else O1)
    assert rel_err(O, O_ref) < 4e-3
    # direct fp32 PyTorch statement of gemma.py:234-272 (rows are (token, head); the mask is per token)
    logits = torch.einsum("bthd,bsd->bths", Q.float(), Kc[:, :T].float())
    logits = torch.where(dense[:, :, None, :], logits, torch.tensor(-2.3819763e38, device=DEV))
    p32 = torch.softmax(logits, -1)
    O32 = torch.einsum("bths,bsd->bthd", p32.bfloat16().float(), Vc[:, :T].float()).reshape(B, R, HD)
    assert rel_err(P[:, :, :T], p32.reshape(B, R, T)) < 3e-3
    assert (P[:, :, :T].float() - p32.reshape(B, R, T)).abs().max() < 2 ** -8 + 1e-6  # half a bf16 ulp at p ~ 1
    assert P[:, :, T:].abs().max() == 0 if Tpad > T else True
    assert rel_err(O, O32) < 4e-3
    # without P output (inference)
    O0b, O1b = torch.zeros_like(O0), torch.zeros_like(O1)
    ops.fa_gemma_fwd(Q, Kc, Vc, bits32, None, O0b, O1b, B, R, NH, Tq, T, Tpad, W32, split, HD)
    assert torch.equal(O0b, O0) and torch.equal(O1b, O1)


def test_fa_pair_variant_is_bit_identical_to_the_product_kernel():
    """`LAPB_FA_PAIR=1` (fa_gemma_pair.cu, cta_group::2) against the product K1 through tools/fa_variant.py in two
    subprocesses (the switch is read once per process): identical checksums of O0, O1 and P on the training shape and on a
    ragged one (T = 333: 128-key last chunk, odd tile count).  First verified on hardware at the end of round 1."""
    import json, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

    def run(pair):
        env = dict(os.environ, LAPB_FA_PAIR=str(pair))
        out = subprocess.run([sys.executable, os.path.join(root, "tools", "fa_variant.py")], env=env, capture_output=True,
                             text=True, timeout=300, cwd=root)
        assert out.returncode == 0, out.stderr[-2000:]
        return json.loads(out.stdout.strip().splitlines()[-1])

    a, b = run(0), run(1)
    for shape in ("train", "ragged"):
        for k in ("O0", "O1", "P", "finite"):
            assert a[shape][k] == b[shape][k], (shape, k)


@pytest.mark.parametrize("Ni,nh,Np,mode", [(3, 16, 256, 0), (2, 2, 64, 0), (2, 2, 16, 1), (1, 16, 256, 1), (2, 4, 729, 0),
                                           (1, 2, 200, 0), (12, 16, 256, 0), (12, 16, 200, 1), (40, 16, 96, 0)])
def test_fused_vit_attention_matches_unfused_and_torch(Ni, nh, Np, mode):
    """K2 (csrc/fa_vit.cu) vs the GEMM + vit_softmax + GEMM path it replaces (same rounding points: equal up to the
    fp32 summation order of the softmax denominator) and vs a direct PyTorch statement of flax's attention with the
    bf16 / fp32 softmax (siglip.py:88-93).  Np = 729 is the 384 px case (3 key chunks, ragged last chunk, no P output:
    Np % 8 != 0), Np = 200 a ragged single chunk, Np = 16 / 64 the test-model sizes; the cases with more than two waves of
    query tiles (12 x 16 heads x 2 tiles, 40 x 16 x 1) run the one-CTA-per-head pair kernel (both tiles / a single tile)."""
    hd = 72
    W = nh * hd
    torch.manual_seed(Np + nh)
    qkv = (torch.randn(Ni * Np, 3 * W, device=DEV) * 0.5).bfloat16()
    qkv[:, :W] = (qkv[:, :W].float() * 0.3).bfloat16()
    want_p = Np % 8 == 0
    O = torch.full((Ni * Np, W), 9.0, device=DEV, dtype=torch.bfloat16)
    P = torch.full((Ni, nh, Np, Np), 7.0, device=DEV, dtype=torch.bfloat16) if want_p else None
    ops.vit_attn_fwd(qkv, O, P, Ni, nh, Np, hd, mode)
    torch.cuda.synchronize()
    q, k, v = (qkv[:, i * W:(i + 1) * W].float().view(Ni, Np, nh, hd) for i in range(3))
    logits = torch.einsum("nqhd,nkhd->nhqk", q, k).bfloat16().float()
    m = logits.max(-1, keepdim=True).values
    if mode == 0:
        e = torch.exp((logits - m).bfloat16().float()).bfloat16().float()
        p_ref = (e / e.sum(-1, keepdim=True).bfloat16().float()).bfloat16()
    else:
        p_ref = torch.softmax(logits, -1).bfloat16()
    o_ref = torch.einsum("nhqk,nkhd->nqhd", p_ref.float(), v).reshape(Ni * Np, W)
    assert torch.isfinite(O.float()).all()
    assert rel_err(O, o_ref) < 4e-3, rel_err(O, o_ref)
    if want_p:
        assert rel_err(P, p_ref) < 3e-3
        assert (P.float() - p_ref.float()).abs().max() <= 2 ** -8 + 1e-6
        # unfused engine path on the same inputs
        Pu = torch.zeros(Ni, nh, Np, Np, device=DEV, dtype=torch.bfloat16)
        Ou = torch.zeros(Ni * Np, W, device=DEV, dtype=torch.bfloat16)
        qf = qkv.view(-1)
        ops.gemm(qf, qf[W:], Pu, M=Np, N=Np, K=hd, lda=3 * W, ldb=3 * W, ldc=Np, batch_i=nh, batch_o=Ni,
                 a_bs=(hd, Np * 3 * W), b_bs=(hd, Np * 3 * W), c_bs=(Np * Np, nh * Np * Np))
        ops.vit_softmax_fwd(Pu, Ni * nh * Np, Np, Np, mode)
        ops.gemm(Pu, qf[2 * W:], Ou, M=Np, N=hd, K=Np, b_major=1, lda=Np, ldb=3 * W, ldc=W, batch_i=nh, batch_o=Ni,
                 a_bs=(Np * Np, nh * Np * Np), b_bs=(hd, Np * 3 * W), c_bs=(hd, Np * W))
        torch.cuda.synchronize()
        # identical except where bf16(sum) sits on a rounding boundary (a whole row then moves by one ulp) and where
        # e * (1 / sum) and e / sum differ in the last fp32 bit right at a bf16 boundary (~3e-5 of the elements)
        assert (P != Pu).float().mean().item() < 2e-3
        assert (P.float() - Pu.float()).abs().max() <= 2 ** -8 + 1e-6
        assert rel_err(O, Ou) < 2e-3
    # without P (inference) the output is the same
    O2 = torch.zeros_like(O)
    ops.vit_attn_fwd(qkv, O2, None, Ni, nh, Np, hd, mode)
    assert torch.equal(O2, O)


@pytest.mark.parametrize("H,W,u8", [(480, 640, True), (200, 300, False), (100, 150, True), (360, 224, False), (224, 224, True)])
def test_image_resize_pad_matches_host_statement(H, W, u8):
    """Device `resize_with_pad` (+ the uint8 -> [-1, 1] conversion) vs `lap_b200.image_tools.resize_with_pad`, the numpy
    statement pinned against the reference's own `resize_with_pad_torch` (model_adapter.py:113-116, image_tools.py:11-52)."""
    from lap_b200 import image_tools
    rng = np.random.default_rng(H + W)
    B, S = 3, 224
    img = rng.integers(0, 256, (B, H, W, 3), dtype=np.uint8) if u8 else rng.uniform(-1, 1, (B, H, W, 3)).astype(np.float32)
    ref = image_tools.resize_with_pad(img, S, S)
    ref = ref.astype(np.float32) / 255.0 * 2.0 - 1.0 if u8 else ref
    plan = image_tools.resize_plan(H, W, S, S)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    out = torch.full((B, S, S, 3), 5.0, device=DEV)
    ops.image_resize_pad(d(img), out, B, H, W, S, S, plan["rh"], plan["rw"], plan["ph0"], plan["pw0"], d(plan["ystart"]),
                         d(plan["yw"]), plan["ytaps"], d(plan["xstart"]), d(plan["xw"]), plan["xtaps"])
    got = out.cpu().numpy()
    diff = np.abs(got - ref)
    if u8:  # a pixel whose filtered value sits on x.5 may round the other way (fp32 summation order): one level of 255
        assert diff.max() <= 2.0 / 255.0 + 1e-6 and (diff > 1e-6).mean() < 2e-3
    else:
        assert diff.max() < 2e-5


@pytest.mark.parametrize("u8", [False, True])
def test_image_augment_matches_numpy_statement(u8):
    """Device augmentation (crop 95 % -> resize -> rotate as one bilinear resampling, then brightness / contrast /
    saturation; explicit per-sample parameters) vs oracle/image_oracle.py (model_adapter.py:118-151)."""
    from oracle import image_oracle as IO
    rng = np.random.default_rng(3)
    B, S = 5, 224
    raw = rng.integers(0, 256, (B, S, S, 3), dtype=np.uint8) if u8 else rng.uniform(-1, 1, (B, S, S, 3)).astype(np.float32)
    as_float = raw.astype(np.float32) / 255.0 * 2.0 - 1.0 if u8 else raw
    p = IO.draw_params(rng, B, S, S, skip=[0, 0, 1, 0, 0])
    p[0, 2], p[1, 2] = 5.0, -5.0           # extreme angles: corners sample outside the image (zero fill)
    p[3, 3:6] = (0.2, -0.2, 0.2)
    p[4, 3:6] = (-0.2, 0.2, -0.2)
    ref = IO.augment(as_float, p)
    p8 = np.zeros((B, 8), np.float32)
    p8[:, :7] = p
    out = torch.zeros((B, S, S, 3), device=DEV)
    ops.image_augment(torch.from_numpy(raw).to(DEV), out, B, S, S, torch.from_numpy(p8).to(DEV))
    got = out.cpu().numpy()
    assert np.abs(got - ref).max() < 2e-4, np.abs(got - ref).max()     # fp32 coordinate arithmetic, 224-pixel lever arm
    assert np.array_equal(got[2], as_float[2])                         # skipped (VQA) sample: untouched
    assert np.abs(got[0] - as_float[0]).mean() > 0.05                  # the others really moved


@pytest.mark.parametrize("Bn,R,T,hd", [(2, 1408, 704, 256), (3, 200, 192, 72)])
def test_softmax_backward_fused_into_the_dp_gemm(Bn, R, T, hd):
    """dS = P o (dP - rowsum(P o dP)), produced by the dP = dO V^T GEMM's epilogue (LAPB_EPI_SOFTMAX_BWD) with the row
    term from lapb200_rowdot(dO, O) — against autograd through an fp32 softmax and against the separate softmax_bwd pass."""
    torch.manual_seed(R + hd)
    logits = torch.randn(Bn, R, T, device=DEV) * 2
    logits[:, :, T - 7:] = -1e30                       # masked keys: P = 0 there
    P = torch.softmax(logits, -1).bfloat16()
    V = torch.randn(Bn, T, hd, device=DEV).bfloat16()
    dO = (torch.randn(Bn, R, hd, device=DEV) * 0.1).bfloat16()
    O = torch.bmm(P.float(), V.float()).bfloat16()
    delta = torch.zeros(Bn, R, device=DEV)
    ops.rowdot(dO, O, delta, R, hd, hd, hd, nbi=Bn, d_bs=(R * hd, 0), o_bs=(R * hd, 0))
    assert rel_err(delta, (dO.float() * O.float()).sum(-1)) < 1e-5
    dS = torch.zeros(Bn, R, T, device=DEV, dtype=torch.bfloat16)
    ops.gemm(dO, V, dS, M=R, N=T, K=hd, ldc=T, batch_i=Bn, a_bs=(R * hd, 0), b_bs=(T * hd, 0), c_bs=(R * T, 0),
             epi=ops.EPI_SOFTMAX_BWD, C2=P, ldc2=T, bias=delta)
    # separate pass (what the fused epilogue replaces)
    dP = torch.zeros_like(dS)
    ops.gemm(dO, V, dP, M=R, N=T, K=hd, ldc=T, batch_i=Bn, a_bs=(R * hd, 0), b_bs=(T * hd, 0), c_bs=(R * T, 0))
    dS2 = torch.zeros_like(dS)
    ops.softmax_bwd(P, dP, dS2, Bn * R, T)
    # fp32 statement
    dP32 = torch.bmm(dO.float(), V.float().transpose(1, 2))
    p32 = P.float()
    ref = p32 * (dP32 - (p32 * dP32).sum(-1, keepdim=True))
    assert rel_err(dS, ref) < 8e-3, rel_err(dS, ref)
    assert rel_err(dS2, ref) < 8e-3
    assert rel_err(dS, dS2) < 8e-3
    assert dS[:, :, T - 7:].abs().max() == 0


@pytest.mark.parametrize("M,N,K,ks,ln,cg,bn", [(692, 2048, 16384, 3, False, 2, 256), (512, 1152, 4304, 4, True, 1, 128),
                                               (100, 256, 1024, 2, False, 1, 128)])
def test_slab_split_k_with_fused_residual_norm(M, N, K, ks, ln, cg, bn):
    """Batch-1 prefix pass: a small-M / long-K projection as `ks` deterministic split-K slabs (plain stores, no atomics)
    whose sum + bias + residual + the next block's norm happen in ONE kernel (lapb200_resid_norm_fwd) — against the
    single-pass GEMM with the residual epilogue followed by the stand-alone norm kernel (same rounding points: equal up to
    the fp32 summation order), and bit-identical from run to run."""
    torch.manual_seed(M + K)
    A = (torch.randn(M, K, device=DEV) * 0.1).bfloat16()
    W = (torch.randn(N, K, device=DEV) * 0.05).bfloat16()
    R = torch.randn(M, N, device=DEV).bfloat16()
    bias = torch.randn(N, device=DEV) * 0.1 if ln else None
    scale = torch.randn(N, device=DEV) * 0.2
    nb = torch.randn(N, device=DEV) * 0.1
    # reference path
    X2 = torch.zeros(M, N, device=DEV, dtype=torch.bfloat16)
    ops.gemm(A, W, X2, M=M, N=N, K=K, epi=ops.EPI_RESID, resid=R, bias=bias)
    Y = torch.zeros_like(X2)
    rstd = torch.zeros(M, device=DEV)
    mean = torch.zeros(M, device=DEV)
    if ln:
        ops.layernorm_fwd(X2, scale, nb, Y, mean, rstd, M, N)
    else:
        ops.rmsnorm_fwd(X2, Y, rstd, M, N, scale=scale)
    # slab path
    def run():
        acc = torch.full((ks, M, N), 7.0, device=DEV)          # stale contents must not matter: slabs are overwritten
        ops.gemm(A, W, acc, M=M, N=N, K=K, k_splits=ks, split_stride=M * N, cta_group=cg, block_n=bn)
        x2 = torch.zeros(M, N, device=DEV, dtype=torch.bfloat16)
        y = torch.zeros_like(x2)
        r2, m2 = torch.zeros(M, device=DEV), torch.zeros(M, device=DEV)
        ops.resid_norm_fwd(R, acc, ks, M * N, bias, x2, ln, scale, nb if ln else None, y, m2 if ln else None, r2, M, N)
        return acc, x2, y, r2
    acc, x2, y, r2 = run()
    ref32 = A.float() @ W.float().T
    assert rel_err(acc.sum(0), ref32) < 1e-5
    assert (x2 != X2).float().mean().item() < 2e-3          # a different fp32 summation order flips a rare bf16 rounding
    assert rel_err(x2, X2) < 1e-3 and rel_err(y, Y) < 2e-3 and rel_err(r2, rstd) < 1e-4
    acc_b, x2_b, y_b, _ = run()
    assert torch.equal(acc, acc_b) and torch.equal(x2, x2_b) and torch.equal(y, y_b)
    # finalisation only (y = None)
    x3 = torch.zeros_like(x2)
    ops.resid_norm_fwd(R, acc, ks, M * N, bias, x3, ln, None, None, None, None, None, M, N)
    assert torch.equal(x3, x2)
