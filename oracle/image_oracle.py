"""ORACLE (test infrastructure, not product code) — numpy restatement of the image side of `preprocess_observation`
(src/lap/models/model_adapter.py:83-181):

  * `resize_with_pad` (model_adapter.py:113-116 -> OP/shared/image_tools.py:11-52) is NOT here: its host-side numpy
    statement `lap_b200.image_tools.resize_with_pad` (pinned against the reference's own `resize_with_pad_torch`) is what
    the device kernel is checked against.
  * train-time augmentation (model_adapter.py:118-151): `augmax.Chain(RandomCrop(0.95 W, 0.95 H), Resize(W, H),
    Rotate((-5, 5)), ColorJitter(brightness=0.2, contrast=0.2, saturation=0.2))` on the image mapped to [0, 1], skipped for
    VQA samples, mapped back to [-1, 1].

PARITY UNPINNED for the augmentation: `augmax` is a third-party dependency that is absent from the reference tree, from
this image and from the wheelhouse (pyproject.toml lists it unpinned), and its random draws come from jax's threefry
stream.  What is restated here is the STRUCTURE the reference composes — one geometric resampling (augmax chains
geometric transforms lazily into a single coordinate map sampled once, bilinearly) followed by the colour jitter — with
every random quantity an EXPLICIT parameter, exactly like `noise=` / `time=` on compute_loss:

    params[b] = (crop_y, crop_x, angle_deg, brightness, contrast, saturation, skip)

  crop_y / crop_x  top-left corner (pixels, may be fractional) of the int(0.95 H) x int(0.95 W) crop window
  angle_deg        rotation of the image content about the image centre, counter-clockwise positive
  brightness       b in [-0.2, 0.2]: v <- v (1 + b) for b < 0, v (1 - b) + b for b >= 0   (blend towards black / white)
  contrast         c in [-0.2, 0.2]: v <- (v - 0.5) (1 + c) + 0.5                         (about mid-grey)
  saturation       s in [-0.2, 0.2]: v <- g + (v - g) (1 + s), g = 0.299 R + 0.587 G + 0.114 B
  skip             1: the sample passes through unchanged (the reference's `vqa_mask`)

Output pixel (y, x) samples the input at  crop + (R(angle) ((y, x) - centre) + centre + 0.5) * (crop_size / size) - 0.5
with bilinear weights; taps outside the image contribute 0 (black in [0, 1] space).  The result is clipped to [0, 1].
The colour formulas are the common definitions; augmax's exact contrast curve could not be checked offline.
"""
from __future__ import annotations

import numpy as np

CROP_FRACTION = 0.95
LUMA = np.array([0.299, 0.587, 0.114], dtype=np.float32)


def augment(images: np.ndarray, params: np.ndarray) -> np.ndarray:
    """images [B, H, W, 3] float32 in [-1, 1]; params [B, 7] float32 (see module docstring) -> same shape / range."""
    images = np.asarray(images, dtype=np.float32)
    params = np.asarray(params, dtype=np.float32)
    B, H, W, C = images.shape
    ch, cw = int(H * CROP_FRACTION), int(W * CROP_FRACTION)
    out = np.empty_like(images)
    yy, xx = np.meshgrid(np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing="ij")
    cy0, cx0 = np.float32((H - 1) / 2.0), np.float32((W - 1) / 2.0)
    for b in range(B):
        crop_y, crop_x, angle, br, co, sa, skip = (np.float32(v) for v in params[b])
        if skip > 0.5:
            out[b] = images[b]
            continue
        img = images[b] * np.float32(0.5) + np.float32(0.5)
        th = np.float32(np.deg2rad(np.float64(angle)))
        cs, sn = np.float32(np.cos(np.float64(th))), np.float32(np.sin(np.float64(th)))
        dy, dx = yy - cy0, xx - cx0
        # content rotated counter-clockwise by `angle`  <=>  output pixel reads from the position rotated clockwise
        ry = cs * dy - sn * dx + cy0
        rx = sn * dy + cs * dx + cx0
        sy = (ry + np.float32(0.5)) * np.float32(ch / H) - np.float32(0.5) + crop_y
        sx = (rx + np.float32(0.5)) * np.float32(cw / W) - np.float32(0.5) + crop_x
        y0, x0 = np.floor(sy), np.floor(sx)
        fy, fx = sy - y0, sx - x0
        y0, x0 = y0.astype(np.int64), x0.astype(np.int64)
        acc = np.zeros((H, W, C), dtype=np.float32)
        for oy, wy in ((0, 1.0 - fy), (1, fy)):
            for ox, wx in ((0, 1.0 - fx), (1, fx)):
                iy, ix = y0 + oy, x0 + ox
                ok = (iy >= 0) & (iy < H) & (ix >= 0) & (ix < W)
                tap = img[np.clip(iy, 0, H - 1), np.clip(ix, 0, W - 1)]
                acc += np.where(ok[..., None], tap, np.float32(0.0)) * (wy * wx).astype(np.float32)[..., None]
        v = acc
        v = v * (1 + br) if br < 0 else v * (1 - br) + br
        v = (v - np.float32(0.5)) * (1 + co) + np.float32(0.5)
        g = (v * LUMA).sum(-1, keepdims=True)
        v = g + (v - g) * (1 + sa)
        out[b] = np.clip(v, 0.0, 1.0) * np.float32(2.0) - np.float32(1.0)
    return out


def draw_params(rng: np.random.Generator, B: int, H: int, W: int, skip=None) -> np.ndarray:
    """One set of augmentation parameters per sample with the reference's ranges (model_adapter.py:127-141)."""
    ch, cw = int(H * CROP_FRACTION), int(W * CROP_FRACTION)
    p = np.zeros((B, 7), dtype=np.float32)
    p[:, 0] = rng.uniform(0, H - ch, B)
    p[:, 1] = rng.uniform(0, W - cw, B)
    p[:, 2] = rng.uniform(-5, 5, B)
    p[:, 3:6] = rng.uniform(-0.2, 0.2, (B, 3))
    if skip is not None:
        p[:, 6] = np.asarray(skip, dtype=np.float32)
    return p
