"""Synthetic RLDS-shaped batches (modelled on FakeDataset, OP/training/data_loader.py:99-127, but with
meaningful masks).  Host-side numpy; the batch is what `RLDSDataLoader` would yield before `_to_device`
(src/lap/datasets/data_loader.py:235-246,327): a dict in `Observation.from_dict` form plus `actions`.
"""
from __future__ import annotations

import numpy as np

from .config import LAPConfig


def synthetic_batch(cfg: LAPConfig, batch_size: int, *, step: int = 0, rank: int = 0, uint8_images: bool = False,
                    with_langact: bool = True) -> dict:
    rng = np.random.default_rng(1000 * step + rank)
    B, L, S = batch_size, cfg.max_token_len, cfg.image_size
    images, image_masks = {}, {}
    for i, key in enumerate(cfg.image_keys):
        if uint8_images:
            img = rng.integers(0, 256, size=(B, S, S, 3), dtype=np.uint8)
        else:
            img = rng.uniform(-1.0, 1.0, size=(B, S, S, 3)).astype(np.float32)
        mask = np.ones(B, dtype=bool)
        if i > 0:  # 10 % of wrist cameras dropped: mask False, pixels -1 (uint8 zeros)
            drop = rng.random(B) < 0.1
            mask &= ~drop
            img[drop] = 0 if uint8_images else -1.0
        images[key], image_masks[key] = img, mask
    tokens = np.zeros((B, L), dtype=np.int32)
    prompt_mask = np.zeros((B, L), dtype=bool)
    langact_mask = np.zeros((B, L), dtype=bool)
    lo_p, hi_p = max(2, int(L * 0.22)), max(3, int(L * 0.36))
    lo_l, hi_l = max(2, int(L * 0.09)), max(3, int(L * 0.22))
    for b in range(B):
        n_p = int(rng.integers(lo_p, hi_p + 1))
        n_l = int(rng.integers(lo_l, hi_l + 1))
        n_l = min(n_l, L - n_p)
        ids = rng.integers(3, cfg.vocab_size, size=n_p + n_l)
        ids[0] = 2  # BOS
        ids[-1] = 1  # EOS (lap.py:32)
        tokens[b, : n_p + n_l] = ids
        prompt_mask[b, : n_p + n_l] = True
        langact_mask[b, n_p : n_p + n_l] = True
    batch = {
        "image": images,
        "image_mask": image_masks,
        "state": rng.uniform(-1, 1, size=(B, cfg.action_dim)).astype(np.float32),
        "tokenized_prompt": tokens,
        "tokenized_prompt_mask": prompt_mask,
        "token_loss_mask": np.ones((B, L), dtype=bool),
        "sample_mask": rng.random(B) < 0.9,
        "actions": rng.uniform(-1, 1, size=(B, cfg.action_horizon, cfg.action_dim)).astype(np.float32),
        # explicit flow-matching randomness (lap.py:193-194), so results do not depend on an RNG stream
        "noise": rng.standard_normal(size=(B, cfg.action_horizon, cfg.action_dim)).astype(np.float32),
        "time": (rng.beta(1.5, 1.0, size=B) * 0.999 + 0.001).astype(np.float32),
    }
    if with_langact:
        batch["tokenized_langact_mask"] = langact_mask
    return batch
