"""Reading / writing the Orbax `params` checkpoint item in its plain-directory layout.

Reference: `restore_params` (third_party/openpi/src/openpi/models/model.py:286-332: PyTreeCheckpointer, item `params`, the
`value` suffix of nnx.State stripped) and `_split_params` (src/lap/training/checkpoints.py:529-547: the `params` item holds
the EMA weights when EMA is on).

Orbax (the reference pins orbax-checkpoint 0.11.13; it is not installable here) writes a PyTree item in one of two
container layouts, both holding one zarr-v2 array per leaf named by the '.'-joined key path:

  * plain directories (`use_ocdbt=False`):  <item>/<key.path>/.zarray + chunk files "i.j.k" (zstd or raw), plus the
    `_METADATA` / `_sharding` JSON files.  THIS module reads and writes that layout with numpy and the system libzstd.
  * OCDBT (`use_ocdbt=True`, the default and what the released LAP-3B / openpi checkpoints use): the same zarr arrays inside
    tensorstore's B+tree key-value store (`manifest.ocdbt`, `d/*`, `ocdbt.process_*`).  Parsing that store needs tensorstore;
    `read_params` recognises it and raises with the one-line conversion to run where orbax is installed
    (`tools/convert_orbax_checkpoint.py`: restore + re-save with `use_ocdbt=False`, or straight to safetensors).

The writer exists so that (i) the round trip is testable offline and (ii) weights trained here can be handed back to the
reference: `PyTreeCheckpointer().restore` auto-detects the plain layout.  The `_METADATA` file follows orbax 0.11's
`tree_metadata` schema as published; it could not be validated against an orbax install in this image.
"""
from __future__ import annotations

import ctypes
import ctypes.util
import json
import os
from pathlib import Path

import numpy as np

_ZSTD = None


def _zstd():
    global _ZSTD
    if _ZSTD is None:
        name = ctypes.util.find_library("zstd")
        if name is None:
            raise RuntimeError("libzstd not found: zstd-compressed zarr chunks cannot be decoded")
        lib = ctypes.CDLL(name)
        lib.ZSTD_getFrameContentSize.restype = ctypes.c_ulonglong
        lib.ZSTD_getFrameContentSize.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
        lib.ZSTD_decompress.restype = ctypes.c_size_t
        lib.ZSTD_decompress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t]
        lib.ZSTD_compress.restype = ctypes.c_size_t
        lib.ZSTD_compress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
        lib.ZSTD_compressBound.restype = ctypes.c_size_t
        lib.ZSTD_compressBound.argtypes = [ctypes.c_size_t]
        lib.ZSTD_isError.restype = ctypes.c_uint
        lib.ZSTD_isError.argtypes = [ctypes.c_size_t]
        _ZSTD = lib
    return _ZSTD


def zstd_decompress(data: bytes, expected: int | None = None) -> bytes:
    lib = _zstd()
    n = lib.ZSTD_getFrameContentSize(data, len(data))
    if n in (2 ** 64 - 1, 2 ** 64 - 2):  # unknown / error
        if expected is None:
            raise ValueError("zstd frame without content size and no expected size")
        n = expected
    out = ctypes.create_string_buffer(max(int(n), 1))
    got = lib.ZSTD_decompress(out, int(n), data, len(data))
    if lib.ZSTD_isError(got):
        raise ValueError("zstd: corrupt chunk")
    return out.raw[:got]


def zstd_compress(data: bytes, level: int = 1) -> bytes:
    lib = _zstd()
    cap = lib.ZSTD_compressBound(len(data))
    out = ctypes.create_string_buffer(cap)
    got = lib.ZSTD_compress(out, cap, data, len(data), level)
    if lib.ZSTD_isError(got):
        raise ValueError("zstd: compression failed")
    return out.raw[:got]


def _np_dtype(z: str):
    if z == "bfloat16":
        import ml_dtypes
        return np.dtype(ml_dtypes.bfloat16)
    return np.dtype(z)


def _zarr_dtype(dt: np.dtype) -> str:
    if dt.name == "bfloat16":
        return "bfloat16"
    return dt.str if dt.itemsize > 1 else dt.str.replace("<", "|").replace(">", "|")


def read_zarr_array(path: str | os.PathLike) -> np.ndarray:
    """One zarr-v2 array directory (`.zarray` + chunk files) -> numpy.  Supports C / F order, zstd or no compressor, '.' or
    '/' chunk-key separators, missing chunks (= fill_value) and ragged edge chunks; no filters."""
    path = Path(path)
    meta = json.loads((path / ".zarray").read_text())
    if meta.get("zarr_format") != 2:
        raise ValueError(f"{path}: zarr_format {meta.get('zarr_format')} is not supported (expected 2)")
    if meta.get("filters"):
        raise ValueError(f"{path}: zarr filters are not supported")
    comp = meta.get("compressor")
    if comp is not None and comp.get("id") != "zstd":
        raise ValueError(f"{path}: compressor {comp.get('id')!r} is not supported (zstd or none)")
    shape, chunks = tuple(meta["shape"]), tuple(meta["chunks"])
    dt = _np_dtype(meta["dtype"])
    order = meta.get("order", "C")
    sep = meta.get("dimension_separator", ".")
    fill = meta.get("fill_value")
    out = np.zeros(shape, dtype=dt) if fill in (None, 0, 0.0) else np.full(shape, fill, dtype=dt)
    grid = [(-(-s // c)) for s, c in zip(shape, chunks)] or []
    n_el = int(np.prod(chunks)) if chunks else 1
    for idx in np.ndindex(*grid) if grid else [()]:
        key = sep.join(str(i) for i in idx) if idx else "0"
        f = path / key
        if not f.exists():
            continue
        raw = f.read_bytes()
        if comp is not None:
            raw = zstd_decompress(raw, n_el * dt.itemsize)
        block = np.frombuffer(raw, dtype=dt, count=n_el).reshape(chunks if chunks else (), order=order)
        sl = tuple(slice(i * c, min((i + 1) * c, s)) for i, c, s in zip(idx, chunks, shape))
        out[sl] = block[tuple(slice(0, s.stop - s.start) for s in sl)]
    return out


def write_zarr_array(path: str | os.PathLike, a: np.ndarray, *, chunks: tuple | None = None, compress: bool = True) -> None:
    path = Path(path)
    path.mkdir(parents=True, exist_ok=True)
    a = np.asarray(a)
    chunks = tuple(chunks) if chunks is not None else tuple(a.shape)
    chunks = tuple(max(int(c), 1) for c in chunks)
    meta = {"chunks": list(chunks), "compressor": {"id": "zstd", "level": 1} if compress else None,
            "dimension_separator": ".", "dtype": _zarr_dtype(a.dtype), "fill_value": None, "filters": None, "order": "C",
            "shape": list(a.shape), "zarr_format": 2}
    (path / ".zarray").write_text(json.dumps(meta, indent=1))
    grid = [(-(-s // c)) for s, c in zip(a.shape, chunks)]
    for idx in np.ndindex(*grid) if grid else [()]:
        block = np.zeros(chunks, dtype=a.dtype)
        sl = tuple(slice(i * c, min((i + 1) * c, s)) for i, c, s in zip(idx, chunks, a.shape))
        block[tuple(slice(0, s.stop - s.start) for s in sl)] = a[sl]
        raw = np.ascontiguousarray(block).tobytes()
        key = ".".join(str(i) for i in idx) if idx else "0"
        (path / key).write_bytes(zstd_compress(raw) if compress else raw)


def is_ocdbt(item_dir: str | os.PathLike) -> bool:
    d = Path(item_dir)
    return (d / "manifest.ocdbt").exists() or any(d.glob("ocdbt.process_*"))


def read_params(item_dir: str | os.PathLike) -> dict:
    """The reference's `restore_params(dir/"params")`: -> nested "pure dict" {PaliGemma: {...}, action_in_proj: {...}, ...}
    of numpy arrays.  The leading `params` level of the saved item and the trailing `value` of nnx.State key paths are
    stripped exactly as model.py:318-332 does."""
    d = Path(item_dir)
    if not d.is_dir():
        raise FileNotFoundError(d)
    if is_ocdbt(d):
        raise NotImplementedError(
            f"{d} is an OCDBT checkpoint (manifest.ocdbt): tensorstore's B+tree store cannot be parsed without tensorstore. "
            "Convert it once where orbax is installed:  python tools/convert_orbax_checkpoint.py <ckpt>/params <out_dir>")
    flat: dict[tuple, np.ndarray] = {}
    for sub in sorted(p for p in d.iterdir() if p.is_dir() and (p / ".zarray").exists()):
        flat[tuple(sub.name.split("."))] = read_zarr_array(sub)
    if not flat:
        raise ValueError(f"{d}: no zarr arrays found (expected <key.path>/.zarray directories)")
    if all(k[0] == "params" for k in flat):
        flat = {k[1:]: v for k, v in flat.items()}
    if all(k[-1] == "value" for k in flat):
        flat = {k[:-1]: v for k, v in flat.items()}
    tree: dict = {}
    for k, v in flat.items():
        node = tree
        for part in k[:-1]:
            node = node.setdefault(part, {})
        node[k[-1]] = v
    return tree


def write_params(item_dir: str | os.PathLike, tree: dict, *, value_suffix: bool = False, compress: bool = True,
                 max_chunk_bytes: int = 256 << 20) -> None:
    """Write a nested parameter tree (or a '/'-joined flat dict) as an Orbax `params` item in the plain-directory layout:
    one zarr-v2 array per leaf under `params.<key.path>[.value]`, `_METADATA` with the tree structure.  `value_suffix`
    reproduces checkpoints written by the reference's training loop (nnx.State leaves)."""
    d = Path(item_dir)
    d.mkdir(parents=True, exist_ok=True)

    def walk(node, prefix):
        for k, v in node.items():
            if isinstance(v, dict):
                yield from walk(v, prefix + (k,))
            else:
                yield prefix + tuple(str(k).split("/")), v

    meta = {}
    for key, leaf in walk(tree, ()):
        a = leaf.detach().cpu().numpy() if hasattr(leaf, "detach") else np.asarray(leaf)
        kp = ("params",) + key + (("value",) if value_suffix else ())
        # chunk the leading axis so that no chunk exceeds max_chunk_bytes (orbax chooses chunk shapes the same way)
        chunks = list(a.shape)
        if a.ndim and a.nbytes > max_chunk_bytes:
            per_row = max(a.nbytes // max(a.shape[0], 1), 1)
            chunks[0] = max(1, min(a.shape[0], max_chunk_bytes // per_row))
        write_zarr_array(d / ".".join(kp), a, chunks=tuple(chunks), compress=compress)
        meta[str(kp)] = {"key_metadata": [{"key": p, "key_type": 2} for p in kp],
                         "value_metadata": {"value_type": "jax.Array", "skip_deserialize": False}}
    (d / "_METADATA").write_text(json.dumps({"tree_metadata": meta, "use_zarr3": False}, indent=1))
    (d / "_sharding").write_text("{}")
