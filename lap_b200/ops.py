"""Thin Python wrappers over the C ABI of liblapb200.so (include/lapb200.h).

torch is used only as the owner of device memory and streams: every wrapper takes torch CUDA tensors,
passes raw device pointers + sizes + the current CUDA stream through ctypes, and returns torch tensors.
No wrapper has a CPU path: calling one with a CPU tensor raises.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from ._lib import GemmParams, check

EPI_NONE, EPI_BIAS_GELU, EPI_RESID, EPI_GATED_RESID, EPI_GEGLU, EPI_QSCALE, EPI_GEGLU_BWD, EPI_GELU_BWD, EPI_SOFTMAX_BWD = range(9)

# launch counter: every C-ABI kernel entry increments this (bench.py reports it as gpu_launches)
launch_count = 0
# default CTA budget of the persistent GEMM (0 = every SM).  The trainer lowers it while an NCCL all-reduce must be
# co-resident: the GEMM CTAs otherwise own every SM's registers/shared memory and the collective cannot overlap.
gemm_max_ctas = 0


def lib():
    return _lib.load()


def _stream() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: torch.Tensor | None) -> ctypes.c_void_p | None:
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("lap_b200 ops require CUDA tensors (there is no CPU fallback)")
    return ctypes.c_void_p(t.data_ptr())


def _count(n: int = 1) -> None:
    global launch_count
    launch_count += n


def check_device() -> None:
    check(lib().lapb200_check_device(), "check_device")


def gemm(
    A: torch.Tensor,
    B: torch.Tensor,
    C: torch.Tensor,
    *,
    M: int,
    N: int,
    K: int,
    a_major: int = 0,
    b_major: int = 0,
    lda: int | None = None,
    ldb: int | None = None,
    ldc: int | None = None,
    batch_i: int = 1,
    batch_o: int = 1,
    a_bs: tuple[int, int] = (0, 0),
    b_bs: tuple[int, int] = (0, 0),
    c_bs: tuple[int, int] = (0, 0),
    epi: int = EPI_NONE,
    bias: torch.Tensor | None = None,
    resid: torch.Tensor | None = None,
    ldr: int | None = None,
    r_bs: tuple[int, int] = (0, 0),
    gate: torch.Tensor | None = None,
    ldg: int = 0,
    gate_rows: int = 1,
    C2: torch.Tensor | None = None,
    ldc2: int = 0,
    accumulate: bool = False,
    q_cols: int = 0,
    q_div: float = 1.0,
    block_n: int = 0,
    max_ctas: int = 0,
    cta_group: int = 0,
    k_splits: int = 0,
    split_stride: int = 0,
) -> torch.Tensor:
    """C[b] = epi(A[b] @ B[b]^T). Strides are in elements; pointers are taken at the tensors' data_ptr()."""
    assert A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16
    assert C.dtype in (torch.bfloat16, torch.float32)
    p = GemmParams()
    p.A, p.B = _ptr(A), _ptr(B)
    p.a_major, p.b_major = a_major, b_major
    p.lda = lda if lda is not None else (K if a_major == 0 else M)
    p.ldb = ldb if ldb is not None else (K if b_major == 0 else N)
    p.a_bs_i, p.a_bs_o = a_bs
    p.b_bs_i, p.b_bs_o = b_bs
    p.M, p.N, p.K = M, N, K
    p.batch_i, p.batch_o = batch_i, batch_o
    p.C = _ptr(C)
    p.ldc = ldc if ldc is not None else N
    p.c_bs_i, p.c_bs_o = c_bs
    p.c_fp32 = 1 if C.dtype == torch.float32 else 0
    p.accumulate = 1 if accumulate else 0
    p.epi = epi
    if bias is not None:
        assert bias.dtype == torch.float32
    p.bias = _ptr(bias)
    p.resid = _ptr(resid)
    p.ldr = ldr if ldr is not None else (ldc if ldc is not None else N)
    p.r_bs_i, p.r_bs_o = r_bs
    p.gate = _ptr(gate)
    p.ldg = ldg
    p.gate_rows = gate_rows
    p.C2 = _ptr(C2)
    p.ldc2 = ldc2
    p.q_cols = q_cols
    p.q_div = q_div
    p.block_n = block_n
    p.max_ctas = max_ctas if max_ctas else gemm_max_ctas
    p.cta_group = cta_group
    p.k_splits = k_splits
    p.split_stride = split_stride
    prof = _gemm_prof
    if prof is not None:
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
    check(lib().lapb200_gemm_bf16(ctypes.byref(p), _stream()), "gemm_bf16")
    if prof is not None:
        e1.record()
        prof.append((e0, e1, 2.0 * M * N * K * batch_i * batch_o * (2 if epi == EPI_GEGLU else 1),
                     (M, N, K, batch_i * batch_o, a_major, b_major, epi)))
    _count()
    return C


# optional per-launch CUDA-event timing of every GEMM (bench.py roofline)
_gemm_prof: list | None = None


def gemm_profile_begin() -> None:
    global _gemm_prof
    _gemm_prof = []


def gemm_profile_end() -> tuple[float, float, int, dict]:
    """Returns (algorithmic FLOPs, summed kernel ms, launches, per-shape {key: [flops, ms, launches]}) since
    gemm_profile_begin(); syncs the device.  key = (M, N, K, batches, a_major, b_major, epilogue)."""
    global _gemm_prof
    prof, _gemm_prof = _gemm_prof or [], None
    torch.cuda.synchronize()
    by_shape: dict = {}
    flops = ms = 0.0
    for a, b, f, key in prof:
        dt = a.elapsed_time(b)
        flops += f
        ms += dt
        ent = by_shape.setdefault(key, [0.0, 0.0, 0])
        ent[0] += f
        ent[1] += dt
        ent[2] += 1
    return flops, ms, len(prof), by_shape


# ---------------------------------------------------------------------------------------------
# generic caller for the "plain" entry points: tensors -> void*, int -> int64, float -> float, + stream
# ---------------------------------------------------------------------------------------------
def _conv(a):
    if a is None:
        return ctypes.c_void_p(0)
    if isinstance(a, torch.Tensor):
        if not a.is_cuda:
            raise RuntimeError("lap_b200 ops require CUDA tensors (there is no CPU fallback)")
        return ctypes.c_void_p(a.data_ptr())
    if isinstance(a, bool):
        return ctypes.c_int64(int(a))
    if isinstance(a, int):
        return ctypes.c_int64(a)
    if isinstance(a, float):
        return ctypes.c_float(a)
    raise TypeError(f"unsupported argument type {type(a)}")


def call(name: str, *args) -> None:
    fn = getattr(lib(), "lapb200_" + name)
    rc = fn(*[_conv(a) for a in args], _stream())
    check(rc, name)
    _count()


def cast_f32_bf16(src, dst):
    call("cast_f32_bf16", src, dst, src.numel())


def split_hi_lo(src, dst, rows, D):
    call("split_hi_lo", src, dst, rows, D)


def patchify(imgs, out, B, C, H, W, ps, out_hi=None, out_lo=None, pk_pad=0):
    is_u8 = imgs[0].dtype == torch.uint8
    p = list(imgs) + [None] * (3 - len(imgs))
    call("patchify", p[0], p[1], p[2], is_u8, out, out_hi, out_lo, pk_pad, B, C, H, W, ps)


def sgemm(A, B, C, M, N, K, sam, sak, sbn, sbk, ldc=None, bias=None, table=None, table_rows=0, accumulate=False):
    """C[m,n] (+)= sum_k A[m*sam+k*sak] * B[n*sbn+k*sbk] (+bias[n]) (+table[m%table_rows, n]); fp32 math."""
    call("sgemm", A, A.dtype == torch.bfloat16, B, B.dtype == torch.bfloat16, C, C.dtype == torch.bfloat16,
         M, N, K, sam, sak, sbn, sbk, ldc if ldc is not None else N, bias, table, table_rows, accumulate)


def layernorm_fwd(x, scale, bias, y, mean, rstd, M, W):
    call("layernorm_fwd", x, scale, bias, y, mean, rstd, M, W)


def layernorm_bwd(dy, x, scale, mean, rstd, dres, dx, dscale, dbias, M, W):
    call("layernorm_bwd", dy, x, scale, mean, rstd, dres, dx, dscale, dbias, M, W)


def rmsnorm_fwd(x, y, rstd, M, D, *, scale=None, mod=None, ldmod=0, rows_per_sample=1, ldx=None, ldy=None,
                row_idx=None, dup=False):
    call("rmsnorm_fwd", x, ldx if ldx is not None else D, row_idx, scale, mod, ldmod, rows_per_sample, y,
         ldy if ldy is not None else D, dup, rstd, M, D)


def rmsnorm_bwd(dy, x, scale, rstd, dres, dx, dscale, M, D, *, lddy=None, ldx=None, row_idx=None):
    call("rmsnorm_bwd", dy, lddy if lddy is not None else D, x, ldx if ldx is not None else D, row_idx, scale, rstd,
         dres, dx, dscale, M, D)


def ada_rmsnorm_bwd(dy, x, mod, ldmod, rstd, dres, dx, dmod, lddmod, B, rows_per_sample, D):
    call("ada_rmsnorm_bwd", dy, x, mod, ldmod, rstd, dres, dx, dmod, lddmod, B, rows_per_sample, D)


def gated_bwd(dxo, y, gate, ldg, dy, dgate, lddg, B, rows_per_sample, D):
    call("gated_bwd", dxo, y, gate, ldg, dy, dgate, lddg, B, rows_per_sample, D)


def rope_fwd(qkv0, qkv1, positions, timescale, Q, Kc, Vc, B, P, A, Tpad, NH, HD, t_begin, qscale):
    call("rope_fwd", qkv0, qkv1, positions, timescale, Q, Kc, Vc, B, P, A, Tpad, NH, HD, t_begin, float(qscale))


def rope_bwd(dQ, dK, dV, positions, timescale, dqkv0, dqkv1, B, P, A, Tpad, NH, HD, qscale):
    call("rope_bwd", dQ, dK, dV, positions, timescale, dqkv0, dqkv1, B, P, A, Tpad, NH, HD, float(qscale))


def geglu_bwd(dact, gu, M, F, write_act=True):
    """dact <- act (only if write_act: the caller may have kept act from the forward), gu <- [dg | du], in place."""
    call("geglu_bwd", dact, gu, M, F, bool(write_act))


def gelu_bwd(dh, pre, n):
    call("gelu_bwd", dh, pre, n)


def swish_fwd(z, y, y_bf16, n):
    call("swish_fwd", z, y, y_bf16, n)


def swish_bwd(z, dy, dy_bf16, dz, n):
    call("swish_bwd", z, dy, dy_bf16, dz, n)


def colsum(X, ldx, out, M, N):
    call("colsum", X, ldx, out, M, N)


def embed_fwd(ids, E, X, B, L, row_off, rows_per_sample, D, scale):
    call("embed_fwd", ids, E, X, B, L, row_off, rows_per_sample, D, float(scale))


def embed_bwd(ids, dX, dE, B, L, row_off, rows_per_sample, D, scale):
    call("embed_bwd", ids, dX, dE, B, L, row_off, rows_per_sample, D, float(scale))


def scatter_rows(d, rows, dX, R, D):
    call("scatter_rows", d, rows, dX, R, D)


def suffix_inputs(actions, noise, time, x_t, u_t, time_emb, B, AD, W):
    call("suffix_inputs", actions, noise, time, x_t, u_t, time_emb, B, AD, W)


def axpy(x, v, dt, n):
    call("axpy", x, v, float(dt), n)


def mask_build(pm, par, pma, sm, sar, bits, positions, B, P, A, W32, row_begin=0, infer_rows=False):
    call("mask_build", pm, par, pma, sm, sar, bits, positions, B, P, A, W32, row_begin, infer_rows)


def mask_expand(bits, dense, rows, S, W32):
    call("mask_expand", bits, dense, rows, S, W32)


def attn_softmax_fwd(S, bits, P, B, rows_per_batch, G, S_len, ld, W32):
    call("attn_softmax_fwd", S, bits, P, B, rows_per_batch, G, S_len, ld, W32)


def softmax_bwd(P, dP, dS, rows, ld):
    call("softmax_bwd", P, dP, dS, rows, ld)


def vit_softmax_fwd(S, rows, n, ld, mode=0):
    call("vit_softmax_fwd", S, rows, n, ld, mode)


def ce_fwd_bwd(logits, ld, targets, weights, nll, dlogits, ldd, R, V):
    call("ce_fwd_bwd", logits, ld, targets, weights, nll, dlogits, ldd, R, V)


def mse_fwd_bwd(v, u, loss, dv, B, AD, gscale):
    call("mse_fwd_bwd", v, u, loss, dv, B, AD, float(gscale))


def weighted_sum(x, w, out, n, alpha=1.0, accumulate=False):
    call("weighted_sum", x, w, out, n, float(alpha), accumulate)


def opt_num_partials() -> int:
    return int(lib().lapb200_opt_num_partials())


def sumsq_partials(x, n, partials):
    call("sumsq_partials", x, n, partials)


def adamw_ema(p, g, m, v, ema, w16, n, gpartials, n_partials, stats, kernel_begin, kernel_end, *, lr, b1, b2, eps, wd,
              bc1, bc2, clip, ema_decay, ema_on, hyper=None):
    call("adamw_ema", p, g, m, v, ema, w16, n, gpartials, n_partials, stats, kernel_begin, kernel_end, float(lr),
         float(b1), float(b2), float(eps), float(wd), float(bc1), float(bc2), float(clip), float(ema_decay), ema_on, hyper)


def sqrt_scalar(buf, src, dst):
    call("sqrt_scalar", buf, src, dst)


def num_sms() -> int:
    return torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count


def skinny_gemm(X, W, Y, *, M, N, K, ldx=None, ldw=None, ldy=None, epi=EPI_NONE, bias=None, resid=None, ldr=None,
                gate=None, ldg=0, gate_rows=1, Y2=None, ldy2=0):
    """Y[M<=16, N] = epi(X[M,K] @ W[N,K]^T): weight-streaming kernel of the denoise loop."""
    call("skinny_gemm", X, ldx if ldx is not None else K, W, ldw if ldw is not None else K, M, N, K, Y,
         ldy if ldy is not None else N, Y.dtype == torch.float32, epi, bias, resid,
         ldr if ldr is not None else (ldy if ldy is not None else N), gate, ldg, gate_rows, Y2, ldy2)


def decode_attn(Q, Kc, Vc, bits, O, B, Tq, NH, HD, S_len, Tpad, W32):
    call("decode_attn", Q, Kc, Vc, bits, O, B, Tq, NH, HD, S_len, Tpad, W32)


def linear_f32(X, W, Y, M, N, K, bias=None):
    """Y[M,N] = X[M,K] @ W[N,K]^T + bias with fp32 weights/accumulation (X fp32 or bf16, Y fp32 or bf16).
    M <= 16 rows take the warp-per-column GEMV kernel, larger M the tiled fp32 GEMM."""
    if M <= 16:
        call("gemv_f32", X, X.dtype == torch.bfloat16, K, W, bias, Y, N, Y.dtype == torch.bfloat16, M, N, K)
    else:
        sgemm(X, W, Y, M, N, K, K, 1, K, 1, ldc=N, bias=bias)


def denoise_supported(B, A, ad, D1, NH, HD, F1, Pn, Tpad, num_steps) -> bool:
    return bool(lib().lapb200_denoise_supported(*(ctypes.c_int64(int(v)) for v in (B, A, ad, D1, NH, HD, F1, Pn, Tpad,
                                                                                  num_steps))))


def transpose_v(Vc, VcT, L, Tpad, TpadK, HD, Pn):
    call("transpose_v", Vc, VcT, L, Tpad, TpadK, HD, Pn)


def pack_tiles(src, dst, rows, cols, row_stride, col_stride, valid_cols, batch=1, src_bs=0, dst_bs=0):
    call("pack_tiles", src, dst, rows, cols, row_stride, col_stride, valid_cols, batch, src_bs, dst_bs)


def denoise_loop(*, ints: dict, dt: float, qscale: float, times, ptrs: dict, strides: dict) -> None:
    """K10: every Euler step of sample_actions in one persistent cooperative kernel (csrc/denoise.cu)."""
    p = _lib.DenoiseParams()
    for k, v in ints.items():
        setattr(p, k, int(v))
    p.dt, p.qscale = float(dt), float(qscale)
    for i, t in enumerate(times):
        p.times[i] = float(t)
    for k, v in ptrs.items():
        if not v.is_cuda:
            raise RuntimeError("lap_b200 ops require CUDA tensors (there is no CPU fallback)")
        setattr(p, k, v.data_ptr())
    for k, v in strides.items():
        setattr(p, k, int(v))
    check(lib().lapb200_denoise_loop(ctypes.byref(p), _stream()), "denoise_loop")
    _count()


def resid_norm_fwd(resid, acc, nsplit, slab_stride, bias, xout, layernorm, scale, nbias, y, mean, rstd, M, D):
    call("resid_norm_fwd", resid, acc, nsplit, slab_stride, bias, xout, bool(layernorm), scale, nbias, y, mean, rstd, M, D)


def rowdot(dO, O, delta, rows, D, ldd, ldo, nbi=1, nbo=1, d_bs=(0, 0), o_bs=(0, 0), out_rows=None, out_off=0):
    """delta[(bo*nbi + bi)*out_rows + out_off + r] = <dO[bo, bi, r, :D], O[bo, bi, r, :D]> (fp32)."""
    call("rowdot", dO, O, delta, rows, D, ldd, ldo, nbi, nbo, d_bs[0], d_bs[1], o_bs[0], o_bs[1],
         rows if out_rows is None else out_rows, out_off)


def image_resize_pad(src, dst, B, Hin, Win, Hout, Wout, rh, rw, ph0, pw0, ystart, yw, ytaps, xstart, xw, xtaps):
    call("image_resize_pad", src, src.dtype == torch.uint8, dst, B, Hin, Win, Hout, Wout, rh, rw, ph0, pw0, ystart, yw,
         ytaps, xstart, xw, xtaps)


def image_augment(src, dst, B, H, W, params):
    call("image_augment", src, src.dtype == torch.uint8, dst, B, H, W, params)


def vit_attn_fwd(qkv, O, P, Ni, nh, Np, hd, mode=0):
    """K2 (csrc/fa_vit.cu): fused SigLIP attention forward; P = None skips the store of the probabilities."""
    call("vit_attn_fwd", qkv, O, P, Ni, nh, Np, hd, mode)


def fa_gemma_fwd(Q, Kc, Vc, bits, P, O0, O1, B, R, G, Tq, S_len, Tpad, W32, split_row, head_dim):
    call("fa_gemma_fwd", Q, Kc, Vc, bits, P, O0, O1, B, R, G, Tq, S_len, Tpad, W32, split_row, head_dim)
