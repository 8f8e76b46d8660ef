"""Seeded configurations, parameters and inputs shared by tests/golden/make_reference_golden.py (which runs the reference's
PyTorch port on them in the build container) and tests/test_reference_golden.py (which runs the oracle / the CUDA engine on
them anywhere).  Pure numpy: nothing here needs /root/reference."""
from __future__ import annotations

import hashlib

import numpy as np

# ------------------------------------------------------------------------------------------------------------------
# the small π0.5-shaped configurations (8 query heads / 1 kv head are hard-coded in gemma_pytorch.py:215)
# ------------------------------------------------------------------------------------------------------------------
VOCAB = 1024
VIS_WIDTH, VIS_DEPTH, VIS_HEADS, VIS_MLP = 144, 2, 2, 200  # = lap_b200 SigLIP test variant "tiny72/14" (head_dim 72)
ACTION_DIM = 32  # pi0_pytorch.py:100-101 hard-codes 32
CASES = {
    # name: (paligemma_variant, expert_variant, batch, action_horizon, max_token_len, seed)
    # (the reference converter needs width == 8 * head_dim for the PaliGemma tower: convert_jax_model_to_pytorch.py:205-212)
    "a": ("pin_a", "pin_a_expert", 2, 10, 24, 11),
    "b": ("pin_b", "pin_b_expert", 1, 5, 16, 12),
}


def lap_config(case: str):
    from lap_b200.config import LAPConfig

    pv, ev, _, horizon, L, _ = CASES[case]
    return LAPConfig(paligemma_variant=pv, action_expert_variant=ev, action_dim=ACTION_DIM, action_horizon=horizon,
                     max_token_len=L, pi05=True, use_bimanual=True, enable_action_training=True,
                     enable_image_augmentation=False, siglip_variant="tiny72/14", vocab_size=VOCAB)


def seeded_reference_params(cfg, seed: int) -> dict[str, np.ndarray]:
    """A non-degenerate parameter tree in the reference's JAX layout (names/shapes: lap_b200.params.reference_shapes =
    SURVEY Appendix B), from numpy's PCG64 so the test can regenerate it bit-for-bit; std chosen so activations stay O(1)."""
    from lap_b200.params import reference_shapes

    rng = np.random.default_rng(seed)
    out = {}
    for name, shape in reference_shapes(cfg).items():
        leaf = name.rsplit("/", 1)[-1]
        if leaf in ("bias",):
            a = 0.05 * rng.standard_normal(shape)
        elif leaf == "scale":
            # RMSNorm scale is used as (1 + scale); LayerNorm scale multiplies directly
            a = (0.1 * rng.standard_normal(shape)) if "llm/" in name else (1.0 + 0.1 * rng.standard_normal(shape))
        elif leaf == "pos_embedding":
            a = 0.1 * rng.standard_normal(shape)
        elif leaf == "input_embedding":
            a = 0.06 * rng.standard_normal(shape)
        else:
            if "Dense_0/kernel" in name and "llm/" in name:
                fan_in = shape[-2]  # adaRMS modulation Dense [L, cond, 3*width]
            elif "q_einsum" in name or "kv_einsum" in name:
                fan_in = shape[-2]  # [.., heads, width, head_dim]
            elif "attn_vec_einsum" in name:
                fan_in = shape[-3] * shape[-2]  # [L, heads, head_dim, width]
            elif "gating_einsum" in name:
                fan_in = shape[-2]  # [L, 2, width, mlp]
            elif "embedding/kernel" in name:
                fan_in = shape[0] * shape[1] * shape[2]  # [ps, ps, 3, width]
            elif "MultiHeadDotProductAttention_0/out/kernel" in name:
                fan_in = shape[-3] * shape[-2]  # [L, heads, hd, width]
            elif "MultiHeadDotProductAttention_0" in name:
                fan_in = shape[-3]  # [L, width, heads, hd]
            else:
                fan_in = shape[-2]
            a = rng.standard_normal(shape) / np.sqrt(fan_in)
        out[name] = np.ascontiguousarray(a.astype(np.float32))
    return out


def seeded_inputs(cfg, batch: int, seed: int) -> dict[str, np.ndarray]:
    rng = np.random.default_rng(seed + 1000)
    L = cfg.max_token_len
    d = {}
    for k in cfg.image_keys:
        d["image/" + k] = rng.uniform(-1, 1, (batch, cfg.image_size, cfg.image_size, 3)).astype(np.float32)
        d["image_mask/" + k] = np.ones((batch,), dtype=bool)
    if batch > 1:
        d["image_mask/" + cfg.image_keys[-1]][-1] = False  # one dropped camera (key padding over 256 image tokens)
    d["tokenized_prompt"] = rng.integers(0, cfg.vocab_size, (batch, L)).astype(np.int32)
    lens = rng.integers(L // 2, L, (batch,))
    d["tokenized_prompt_mask"] = np.arange(L)[None, :] < lens[:, None]
    d["state"] = rng.uniform(-1, 1, (batch, ACTION_DIM)).astype(np.float32)
    d["actions"] = rng.uniform(-1, 1, (batch, cfg.action_horizon, cfg.action_dim)).astype(np.float32)
    d["noise"] = rng.standard_normal((batch, cfg.action_horizon, cfg.action_dim)).astype(np.float32)
    d["time"] = (rng.beta(1.5, 1.0, (batch,)) * 0.999 + 0.001).astype(np.float32)
    return d


def params_digest(params: dict[str, np.ndarray]) -> str:
    h = hashlib.sha256()
    for k in sorted(params):
        h.update(k.encode())
        h.update(np.ascontiguousarray(params[k]).tobytes())
    return h.hexdigest()


def grad_fingerprint(g: np.ndarray) -> np.ndarray:
    """Three numbers per gradient tensor (sum, L2 norm, dot with a fixed +-1 pattern) — what the fixtures keep of the
    reference's autograd gradients."""
    g = np.asarray(g, dtype=np.float64).reshape(-1)
    sign = np.where((np.arange(g.size, dtype=np.int64) * 2654435761 % 4294967296) & 0x10000, 1.0, -1.0)
    return np.array([g.sum(), np.sqrt((g * g).sum()), (g * sign).sum()], dtype=np.float64)


def pack_rows(x: np.ndarray, stride: int) -> np.ndarray:
    """Fixtures keep every `stride`-th token row of the big [B, tokens, width] activations (plus always the last 32 rows)."""
    return x[:, row_index(x.shape[1], stride)]


def row_index(n: int, stride: int) -> np.ndarray:
    return np.unique(np.concatenate([np.arange(0, n, stride), np.arange(max(0, n - 32), n)]))


# ------------------------------------------------------------------------------------------------------------------
# LAP-shaped cases (two cameras, lang-action tokens, loss/sample masks) for make_reference_lap_golden.py
# ------------------------------------------------------------------------------------------------------------------
LAP_CASES = {
    "a": dict(gemma="pin_a", expert="pin_a_expert", batch=3, horizon=10, L=40, seed=21, lang_w=0.4, act_w=1.0),
    "b": dict(gemma="pin_b", expert="pin_b_expert", batch=2, horizon=6, L=28, seed=22, lang_w=1.0, act_w=1.0),
}


def lap_case_config(case: str):
    from lap_b200.config import LAPConfig

    c = LAP_CASES[case]
    return LAPConfig(paligemma_variant=c["gemma"], action_expert_variant=c["expert"], action_dim=ACTION_DIM,
                     action_horizon=c["horizon"], max_token_len=c["L"], pi05=True, use_bimanual=False,
                     enable_action_training=True, enable_langact_training=True, enable_image_augmentation=False,
                     language_loss_weight=c["lang_w"], action_loss_weight=c["act_w"], siglip_variant="tiny72/14",
                     vocab_size=VOCAB)


def lap_case_inputs(cfg, batch: int, seed: int) -> dict[str, np.ndarray]:
    """RLDS-shaped batch: prompt tokens, then a lang-action span ending in EOS, then padding; one sample
    masked out of the language loss, one wrist camera dropped, a few loss-mask holes."""
    d = seeded_inputs(cfg, batch, seed)
    rng = np.random.default_rng(seed + 2000)
    L = cfg.max_token_len
    n_p = rng.integers(L // 4, L // 2, (batch,))
    n_l = rng.integers(4, L // 3, (batch,))
    pos = np.arange(L)[None, :]
    d["tokenized_prompt_mask"] = pos < (n_p + n_l)[:, None]
    d["tokenized_langact_mask"] = (pos >= n_p[:, None]) & (pos < (n_p + n_l)[:, None])
    tok = rng.integers(3, cfg.vocab_size, (batch, L)).astype(np.int32)
    tok[:, 0] = 2
    tok[np.arange(batch), n_p + n_l - 1] = 1
    tok[~d["tokenized_prompt_mask"]] = 0
    d["tokenized_prompt"] = tok
    d["token_loss_mask"] = rng.random((batch, L)) < 0.9
    d["sample_mask"] = np.ones((batch,), dtype=bool)
    d["sample_mask"][0] = False
    d["image_mask/left_wrist_0_rgb"][:] = True
    d["image_mask/left_wrist_0_rgb"][-1] = False
    return d
