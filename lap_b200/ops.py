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

EPI_NONE, EPI_BIAS_GELU, EPI_RESID, EPI_GATED_RESID, EPI_GEGLU, EPI_QSCALE = range(6)

# launch counter: every C-ABI kernel entry increments this (bench.py reports it as gpu_launches)
launch_count = 0


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
    p.max_ctas = max_ctas
    check(lib().lapb200_gemm_bf16(ctypes.byref(p), _stream()), "gemm_bf16")
    _count()
    return C
