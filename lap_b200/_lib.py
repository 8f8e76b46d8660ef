"""ctypes binding of liblapb200.so (the sm_100a kernel library) and its in-tree build.

The library is built IN-TREE with nvcc (`-gencode arch=compute_100a,code=sm_100a -lineinfo`), next to the
sources in `lap_b200/csrc/`, so the binary travels with the repo snapshot to the GPU box.  There is no CPU
fallback: importing this module never computes anything, and every kernel call fails loudly (RuntimeError)
if the library is missing or the device is not an sm_100 part.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import threading
from pathlib import Path

CSRC = Path(__file__).resolve().parent / "csrc"
LIB_PATH = CSRC / "liblapb200.so"
INCLUDE = Path(__file__).resolve().parent.parent / "include"
SOURCES = ["api.cu", "gemm.cu", "elementwise.cu", "attention.cu", "loss.cu", "optimizer.cu", "skinny.cu", "fa_gemma.cu", "fa_gemma_pair.cu", "fa_vit.cu", "denoise.cu", "image.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC",
    "-Wno-deprecated-declarations",
]

_lock = threading.Lock()
_lib = None


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def needs_build() -> bool:
    if not LIB_PATH.exists():
        return True
    t = LIB_PATH.stat().st_mtime
    deps = [p for p in CSRC.iterdir() if p.suffix in (".cu", ".cuh", ".h")] + list(INCLUDE.glob("*.h")) + [Path(__file__)]
    return any(p.stat().st_mtime > t for p in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every .cu under csrc/ to objects (in parallel) and link liblapb200.so.

    Safe under torchrun on a fresh checkout: an inter-process file lock serialises the ranks (the first one builds, the
    others find an up-to-date library when they get the lock), objects and the library are written to temporary names
    and moved into place atomically."""
    import fcntl

    if not force and not needs_build():
        return LIB_PATH
    with open(CSRC / ".build.lock", "w") as lockf:
        fcntl.flock(lockf, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():
                return LIB_PATH
            return _build_locked(force, verbose)
        finally:
            fcntl.flock(lockf, fcntl.LOCK_UN)


def _build_locked(force: bool, verbose: bool) -> Path:
    srcs = [CSRC / s for s in SOURCES if (CSRC / s).exists()]
    objs = []
    procs = []
    nvcc = _nvcc()
    hdr_mtime = max(p.stat().st_mtime for p in list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + list(INCLUDE.glob("*.h")))
    for s in srcs:
        o = s.with_suffix(".o")
        objs.append(o)
        if not force and o.exists() and o.stat().st_mtime > max(s.stat().st_mtime, hdr_mtime):
            continue
        tmp = o.with_suffix(f".o.tmp{os.getpid()}")
        cmd = [nvcc, *NVCC_FLAGS, "-I", str(INCLUDE), "-c", str(s), "-o", str(tmp)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((s, o, tmp, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for s, o, tmp, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            tmp.unlink(missing_ok=True)
            raise RuntimeError(f"nvcc failed on {s.name}:\n{out}")
        os.replace(tmp, o)
        if verbose and out.strip():
            print(f"--- {s.name}\n{out}")
    tmp_lib = LIB_PATH.with_suffix(f".so.tmp{os.getpid()}")
    cmd = [nvcc, "-shared", "-o", str(tmp_lib), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        tmp_lib.unlink(missing_ok=True)
        raise RuntimeError(f"link failed:\n{r.stdout}")
    os.replace(tmp_lib, LIB_PATH)
    return LIB_PATH


# ---------------------------------------------------------------------------------------------
# struct mirrors (must match include/lapb200.h)
# ---------------------------------------------------------------------------------------------
c_i32, c_i64, c_f32, c_vp = ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p


class GemmParams(ctypes.Structure):
    _fields_ = [
        ("A", c_vp), ("B", c_vp), ("a_major", c_i32), ("b_major", c_i32),
        ("lda", c_i64), ("ldb", c_i64),
        ("a_bs_i", c_i64), ("a_bs_o", c_i64), ("b_bs_i", c_i64), ("b_bs_o", c_i64),
        ("M", c_i32), ("N", c_i32), ("K", c_i32), ("batch_i", c_i32), ("batch_o", c_i32),
        ("C", c_vp), ("ldc", c_i64), ("c_bs_i", c_i64), ("c_bs_o", c_i64),
        ("c_fp32", c_i32), ("accumulate", c_i32), ("epi", c_i32),
        ("bias", c_vp), ("resid", c_vp), ("ldr", c_i64), ("r_bs_i", c_i64), ("r_bs_o", c_i64),
        ("gate", c_vp), ("ldg", c_i64), ("gate_rows", c_i32),
        ("C2", c_vp), ("ldc2", c_i64), ("q_cols", c_i32), ("q_div", c_f32),
        ("block_n", c_i32), ("max_ctas", c_i32), ("cta_group", c_i32), ("k_splits", c_i32),
        ("reserved0_", c_i32), ("split_stride", c_i64),
    ]


class DenoiseParams(ctypes.Structure):
    """lapb_denoise_params_t (include/lapb200.h)."""
    _fields_ = [
        *[(n, c_i32) for n in ("A", "ad", "D1", "NH", "HD", "F1", "L", "Pn", "Tpad", "TpadK", "W32", "nm", "num_steps")],
        ("dt", c_f32), ("qscale", c_f32), ("times", c_f32 * 16),
        ("qkv_w", c_vp), ("o_w", c_vp), ("gu_w", c_vp), ("down_w", c_vp),
        ("qkv_ls", c_i64), ("o_ls", c_i64), ("gu_ls", c_i64), ("down_ls", c_i64),
        ("mod_w", c_vp), ("mod_b", c_vp),
        *[(n, c_vp) for n in ("ain_w", "ain_b", "tin_w", "tin_b", "tout_w", "tout_b", "aout_w", "aout_b")],
        ("Kc", c_vp), ("VcT", c_vp), ("kc_ls", c_i64), ("vct_ls", c_i64),
        ("bits", c_vp), ("pos", c_vp), ("timescale", c_vp), ("x", c_vp), ("s1", c_vp),
        ("cond16", c_vp), ("mod", c_vp),
        *[(n, c_vp) for n in ("XE", "XE1", "qkv", "O", "act")],
        ("part_o", c_vp), ("part_ml", c_vp), ("sync", c_vp), ("prof", c_vp),
        ("packed", c_i32), ("flags", c_i32),
    ]


def load(build_if_needed: bool = True) -> ctypes.CDLL:
    """Load liblapb200.so (building it first if sources are newer). Raises if it cannot be produced."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if build_if_needed and needs_build():
            build()
        if not LIB_PATH.exists():
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        lib = ctypes.CDLL(str(LIB_PATH))
        lib.lapb200_last_error.restype = ctypes.c_char_p
        lib.lapb200_version.restype = ctypes.c_int
        _lib = lib
        return lib


def last_error() -> str:
    return load().lapb200_last_error().decode()


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise RuntimeError(f"lapb200 {what} failed (code {rc}): {last_error()}")


def exported_symbols() -> list[str]:
    """Every `lapb200_*` symbol declared in include/lapb200.h."""
    import re

    text = (INCLUDE / "lapb200.h").read_text()
    return sorted(set(re.findall(r"\b(lapb200_[a-z0-9_]+)\s*\(", text)))
