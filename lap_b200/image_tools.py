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


def resize_plan(in_h: int, in_w: int, height: int, width: int, *, antialias: bool = True) -> dict:
    """Everything the device kernel `lapb200_image_resize_pad` needs for one (input, output) resolution pair: the size and
    position of the resized region and the two interpolation matrices of `resize_with_pad` in sparse-row form
    (start index + a fixed number of taps per output row, zero-padded weights)."""
    ratio = max(in_w / width, in_h / height)
    rh, rw = int(in_h / ratio), int(in_w / ratio)

    def sparse(w: np.ndarray):
        nz = w != 0
        start = np.where(nz.any(1), nz.argmax(1), 0).astype(np.int32)
        last = np.where(nz.any(1), w.shape[1] - 1 - nz[:, ::-1].argmax(1), 0)
        taps = int((last - start + 1).max())
        idx = start[:, None] + np.arange(taps)[None, :]
        vals = np.where(idx < w.shape[1], np.take_along_axis(w, np.minimum(idx, w.shape[1] - 1), 1), 0.0)
        return start, vals.astype(np.float32), taps

    ys, yw, yt = sparse(_weights(in_h, rh, antialias))
    xs, xw, xt = sparse(_weights(in_w, rw, antialias))
    return dict(rh=rh, rw=rw, ph0=(height - rh) // 2, pw0=(width - rw) // 2, ystart=ys, yw=yw, ytaps=yt, xstart=xs, xw=xw,
                xtaps=xt)


AUG_CROP_FRACTION = 0.95  # model_adapter.py:130,136: RandomCrop(int(w * 0.95), int(h * 0.95))


def draw_augmentation_params(rng: np.random.Generator, batch: int, height: int, width: int, skip=None) -> np.ndarray:
    """Per-sample parameters of the train-time augmentation (model_adapter.py:118-151) with the reference's ranges: crop
    window position uniform over the image, rotation in (-5, 5) degrees, brightness / contrast / saturation in
    (-0.2, 0.2); `skip` marks VQA samples, which the reference leaves untouched.  -> float32 [batch, 8] rows
    (crop_y, crop_x, angle_deg, brightness, contrast, saturation, skip, 0) for `lapb200_image_augment`."""
    ch, cw = int(height * AUG_CROP_FRACTION), int(width * AUG_CROP_FRACTION)
    p = np.zeros((batch, 8), dtype=np.float32)
    p[:, 0] = rng.uniform(0, height - ch, batch)
    p[:, 1] = rng.uniform(0, width - cw, batch)
    p[:, 2] = rng.uniform(-5.0, 5.0, batch)
    p[:, 3:6] = rng.uniform(-0.2, 0.2, (batch, 3))
    if skip is not None:
        p[:, 6] = np.asarray(skip, dtype=np.float32).reshape(batch)
    return p
