"""`resize_with_pad` — the resolution fix-up of `preprocess_observation` (src/lap/models/model_adapter.py:113-116 ->
third_party/openpi/src/openpi/shared/image_tools.py:11-52): aspect-preserving linear resize, clip / round, centre padding
(0 for uint8, -1 for float32 images).  Host-side numpy: images normally arrive at 224 x 224 and this never runs on the hot path.

The resize itself is `jax.image.resize(method="linear")` in the reference — a third-party routine that is absent here.  It is
restated from its published definition (separable weight matrices, half-pixel centres, triangle kernel widened by the
downscale factor = antialiasing, weights renormalised at the borders).  What is pinned: with `antialias=False` the function is
checked against the reference's own `resize_with_pad_torch` (`image_tools.py:55-128`, F.interpolate bilinear) for up- and
down-scaling, and for upscaling both settings coincide; the antialiased DOWNscale is unpinned (tests/test_host.py).
"""
from __future__ import annotations

import numpy as np


def _weights(in_size: int, out_size: int, antialias: bool) -> np.ndarray:
    """[out_size, in_size] interpolation matrix of jax.image.resize (scale.py: compute_weight_mat, triangle kernel)."""
    scale = out_size / in_size
    inv = 1.0 / scale
    kernel_scale = max(inv, 1.0) if antialias else 1.0
    sample = (np.arange(out_size, dtype=np.float64) + 0.5) * inv - 0.5
    x = np.abs(sample[:, None] - np.arange(in_size, dtype=np.float64)[None, :]) / kernel_scale
    w = np.maximum(0.0, 1.0 - x)
    total = w.sum(axis=1, keepdims=True)
    w = np.where(np.abs(total) > 1000.0 * np.finfo(np.float32).eps, w / np.where(total == 0, 1, total), 0.0)
    inside = (sample >= -0.5) & (sample <= in_size - 0.5)
    return (w * inside[:, None]).astype(np.float32)


def resize_with_pad(images, height: int, width: int, *, antialias: bool = True) -> np.ndarray:
    """images [*b, h, w, c] uint8 or float32 in [-1, 1] -> [*b, height, width, c] of the same dtype."""
    images = np.asarray(images)
    if images.dtype not in (np.uint8, np.float32):
        raise ValueError(f"Unsupported image dtype: {images.dtype}")
    batched = images.ndim == 4
    x = images if batched else images[None]
    ch, cw = x.shape[1:3]
    if (ch, cw) == (height, width):
        return images
    ratio = max(cw / width, ch / height)
    rh, rw = int(ch / ratio), int(cw / ratio)
    wh, ww = _weights(ch, rh, antialias), _weights(cw, rw, antialias)
    y = np.einsum("ph,bhwc->bpwc", wh, x.astype(np.float32))
    y = np.einsum("qw,bpwc->bpqc", ww, y)
    if images.dtype == np.uint8:
        y = np.clip(np.round(y), 0, 255).astype(np.uint8)   # round half to even, like jnp.round / torch.round
        pad_value = 0
    else:
        y = np.clip(y, -1.0, 1.0)
        pad_value = -1.0
    ph0, rem_h = divmod(height - rh, 2)
    pw0, rem_w = divmod(width - rw, 2)
    y = np.pad(y, ((0, 0), (ph0, ph0 + rem_h), (pw0, pw0 + rem_w), (0, 0)), constant_values=pad_value)
    return y if batched else y[0]
