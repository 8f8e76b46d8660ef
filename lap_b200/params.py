"""Parameter tree: reference names/shapes (the checkpoint contract) <-> the engine's flat device layout.

Reference tree (names confirmed by third_party/openpi/examples/convert_jax_model_to_pytorch.py
:55-260 and gemma.py:567-574): root keys `PaliGemma/{img,llm}`, `action_in_proj`, `time_mlp_in`, `time_mlp_out`,
`action_out_proj`; leading 18/27 = nn.scan layer axis.  Keys here are '/'-joined paths.

Engine layout: ONE flat fp32 buffer (master params), with identically laid-out flat buffers for grads, Adam mu/nu,
EMA and a bf16 compute copy.  Every GEMM weight is stored [out, in] row-major (K-major B operand of the forward
GEMM); fused projections are stored stacked ([q;k;v], [gate;up], all 37 adaRMS modulation Dense layers).  The GEMM
("kernel") weights come first so that `param_norm` (scripts/train.py:402-415) is one contiguous range.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch

from .config import LAPConfig

ALIGN = 64  # elements; keeps every tensor 256-byte aligned in the fp32 buffer and 128-byte in the bf16 copy


# ---------------------------------------------------------------------------------------------
# reference-layout spec and init
# ---------------------------------------------------------------------------------------------
def reference_shapes(cfg: LAPConfig) -> "OrderedDict[str, tuple[int, ...]]":
    g, e, s = cfg.gemma, cfg.expert, cfg.siglip
    ps, W, Ls = s.patch_size, s.width, s.depth
    nh, hd = s.num_heads, s.head_dim
    L = g.depth
    sh: OrderedDict[str, tuple[int, ...]] = OrderedDict()
    img = "PaliGemma/img/"
    sh[img + "embedding/kernel"] = (ps, ps, 3, W)
    sh[img + "embedding/bias"] = (W,)
    sh[img + "pos_embedding"] = (1, cfg.num_patches, W)
    blk = img + "Transformer/encoderblock/"
    for ln in ("LayerNorm_0", "LayerNorm_1"):
        sh[blk + ln + "/scale"] = (Ls, W)
        sh[blk + ln + "/bias"] = (Ls, W)
    mha = blk + "MultiHeadDotProductAttention_0/"
    for n in ("query", "key", "value"):
        sh[mha + n + "/kernel"] = (Ls, W, nh, hd)
        sh[mha + n + "/bias"] = (Ls, nh, hd)
    sh[mha + "out/kernel"] = (Ls, nh, hd, W)
    sh[mha + "out/bias"] = (Ls, W)
    sh[blk + "MlpBlock_0/Dense_0/kernel"] = (Ls, W, s.mlp_dim)
    sh[blk + "MlpBlock_0/Dense_0/bias"] = (Ls, s.mlp_dim)
    sh[blk + "MlpBlock_0/Dense_1/kernel"] = (Ls, s.mlp_dim, W)
    sh[blk + "MlpBlock_0/Dense_1/bias"] = (Ls, W)
    sh[img + "Transformer/encoder_norm/scale"] = (W,)
    sh[img + "Transformer/encoder_norm/bias"] = (W,)
    sh[img + "head/kernel"] = (W, s.num_classes)
    sh[img + "head/bias"] = (s.num_classes,)
    llm = "PaliGemma/llm/"
    sh[llm + "embedder/input_embedding"] = (cfg.vocab_size, g.width)
    lay = llm + "layers/"
    for i, c in enumerate((g, e)):
        sfx = "" if i == 0 else "_1"
        sh[lay + f"attn/q_einsum{sfx}/w"] = (L, c.num_heads, c.width, c.head_dim)
        sh[lay + f"attn/kv_einsum{sfx}/w"] = (L, 2, c.num_kv_heads, c.width, c.head_dim)
        sh[lay + f"attn/attn_vec_einsum{sfx}/w"] = (L, c.num_heads, c.head_dim, c.width)
        sh[lay + f"mlp{sfx}/gating_einsum"] = (L, 2, c.width, c.mlp_dim)
        sh[lay + f"mlp{sfx}/linear"] = (L, c.mlp_dim, c.width)
        if i == 0 or not cfg.pi05:
            sh[lay + f"pre_attention_norm{sfx}/scale"] = (L, c.width)
            sh[lay + f"pre_ffw_norm{sfx}/scale"] = (L, c.width)
            sh[llm + f"final_norm{sfx}/scale"] = (c.width,)
        else:
            for nm in ("pre_attention_norm", "pre_ffw_norm"):
                sh[lay + f"{nm}{sfx}/Dense_0/kernel"] = (L, c.width, 3 * c.width)
                sh[lay + f"{nm}{sfx}/Dense_0/bias"] = (L, 3 * c.width)
            sh[llm + f"final_norm{sfx}/Dense_0/kernel"] = (c.width, 3 * c.width)
            sh[llm + f"final_norm{sfx}/Dense_0/bias"] = (3 * c.width,)
    D1 = e.width
    sh["action_in_proj/kernel"] = (cfg.action_dim, D1)
    sh["action_in_proj/bias"] = (D1,)
    sh["time_mlp_in/kernel"] = (D1, D1)
    sh["time_mlp_in/bias"] = (D1,)
    sh["time_mlp_out/kernel"] = (D1, D1)
    sh["time_mlp_out/bias"] = (D1,)
    sh["action_out_proj/kernel"] = (D1, cfg.action_dim)
    sh["action_out_proj/bias"] = (cfg.action_dim,)
    return sh


def init_reference_params(cfg: LAPConfig, seed: int = 0, *, reference_zero_init: bool = True) -> dict[str, torch.Tensor]:
    """Random init in the reference layout (CPU fp32).

    Follows the reference initialisers in distribution (lecun-normal fan-in einsums gemma.py:183-199, normal(0.01)
    embedding :144, zeros RMS scale :121 and adaRMS Dense :128; xavier-uniform SigLIP, zero head siglip.py:285;
    nnx.Linear lecun-normal).  `reference_zero_init=False` replaces the zero inits by small random values so tests
    exercise those parameters.  JAX's threefry stream is not reproduced: parity is about the function.
    """
    gen = torch.Generator().manual_seed(seed)
    out: dict[str, torch.Tensor] = {}

    def normal(shape, std):
        return torch.randn(shape, generator=gen, dtype=torch.float32) * std

    for name, shape in reference_shapes(cfg).items():
        leaf = name.rsplit("/", 1)[-1]
        if name.endswith("input_embedding"):
            t = normal(shape, 0.01 if reference_zero_init else 0.05)
        elif "/img/" in name:
            if leaf == "kernel":
                if name.endswith("head/kernel") and reference_zero_init:
                    t = torch.zeros(shape)
                else:
                    if "embedding/kernel" in name:
                        fan_in, fan_out = shape[0] * shape[1] * shape[2], shape[3]
                    elif "out/kernel" in name:
                        fan_in, fan_out = shape[-3] * shape[-2], shape[-1]
                    elif any(k in name for k in ("query", "key", "value")):
                        fan_in, fan_out = shape[-3], shape[-2] * shape[-1]
                    else:
                        fan_in, fan_out = shape[-2], shape[-1]
                    lim = math.sqrt(6.0 / (fan_in + fan_out))
                    t = (torch.rand(shape, generator=gen) * 2 - 1) * lim
            elif leaf == "scale":
                t = torch.ones(shape) if reference_zero_init else 1.0 + normal(shape, 0.1)
            elif leaf == "bias":
                t = torch.zeros(shape) if reference_zero_init else normal(shape, 0.02)
            elif leaf == "pos_embedding":
                t = normal(shape, 1.0 / math.sqrt(shape[-1]))
            else:
                raise KeyError(name)
        elif leaf == "scale":
            t = torch.zeros(shape) if reference_zero_init else normal(shape, 0.1)
        elif "Dense_0" in name:
            t = torch.zeros(shape) if reference_zero_init else normal(shape, 0.02)
        elif leaf == "bias":
            t = torch.zeros(shape) if reference_zero_init else normal(shape, 0.02)
        elif "q_einsum" in name or "kv_einsum" in name:
            t = normal(shape, 1.0 / math.sqrt(shape[-2]))
        elif "attn_vec_einsum" in name:
            t = normal(shape, 1.0 / math.sqrt(shape[-3] * shape[-2]))
        elif leaf in ("gating_einsum", "linear", "kernel"):
            t = normal(shape, 1.0 / math.sqrt(shape[-2]))
        else:
            raise KeyError(name)
        out[name] = t.contiguous()
    return out


# ---------------------------------------------------------------------------------------------
# engine layout
# ---------------------------------------------------------------------------------------------
def n_mod(cfg: LAPConfig) -> int:
    return 2 * cfg.expert.depth + 1  # pre_attention_norm_1, pre_ffw_norm_1 per layer + final_norm_1


def engine_shapes(cfg: LAPConfig) -> tuple["OrderedDict[str, tuple[int, ...]]", list[str]]:
    """Engine tensor shapes in flat order and the list of names counted in param_norm (GEMM kernels)."""
    g, e, s = cfg.gemma, cfg.expert, cfg.siglip
    W, Ls, F = s.width, s.depth, s.mlp_dim
    L = g.depth
    D, D1 = g.width, e.width
    qkv = (g.num_heads + 2 * g.num_kv_heads) * g.head_dim
    assert g.num_kv_heads == 1 and e.num_kv_heads == 1, "engine assumes multi-query attention (kv heads = 1)"
    kern: OrderedDict[str, tuple[int, ...]] = OrderedDict()
    kern["img.patch_w"] = (W, s.patch_size * s.patch_size * 3)
    kern["img.qkv_w"] = (Ls, 3 * W, W)
    kern["img.out_w"] = (Ls, W, W)
    kern["img.fc1_w"] = (Ls, F, W)
    kern["img.fc2_w"] = (Ls, W, F)
    kern["img.head_w"] = (s.num_classes, W)
    kern["g.qkv_w"] = (L, qkv, D)
    kern["g.o_w"] = (L, D, g.num_heads * g.head_dim)
    kern["g.gu_w"] = (L, 2 * g.mlp_dim, D)
    kern["g.down_w"] = (L, D, g.mlp_dim)
    kern["e.qkv_w"] = (L, qkv, D1)
    kern["e.o_w"] = (L, D1, e.num_heads * e.head_dim)
    kern["e.gu_w"] = (L, 2 * e.mlp_dim, D1)
    kern["e.down_w"] = (L, D1, e.mlp_dim)
    kern["e.mod_w"] = (n_mod(cfg), 3 * D1, D1)
    kern["action_in_w"] = (D1, cfg.action_dim)
    kern["time_in_w"] = (D1, D1)
    kern["time_out_w"] = (D1, D1)
    kern["action_out_w"] = (cfg.action_dim, D1)
    other: OrderedDict[str, tuple[int, ...]] = OrderedDict()
    other["g.embed"] = (cfg.vocab_size, D)
    other["img.patch_b"] = (W,)
    other["img.pos"] = (cfg.num_patches, W)
    for nm in ("ln0_s", "ln0_b", "ln1_s", "ln1_b", "out_b", "fc2_b"):
        other["img." + nm] = (Ls, W)
    other["img.qkv_b"] = (Ls, 3 * W)
    other["img.fc1_b"] = (Ls, F)
    other["img.enc_s"] = (W,)
    other["img.enc_b"] = (W,)
    other["img.head_b"] = (s.num_classes,)
    other["g.attn_norm_s"] = (L, D)
    other["g.ffn_norm_s"] = (L, D)
    other["g.final_norm_s"] = (D,)
    other["e.mod_b"] = (n_mod(cfg), 3 * D1)
    other["action_in_b"] = (D1,)
    other["time_in_b"] = (D1,)
    other["time_out_b"] = (D1,)
    other["action_out_b"] = (cfg.action_dim,)
    shapes = OrderedDict(list(kern.items()) + list(other.items()))
    return shapes, list(kern.keys())


class FlatLayout:
    """Offsets of every engine tensor inside the flat buffers."""

    def __init__(self, cfg: LAPConfig):
        self.cfg = cfg
        self.shapes, self.kernel_names = engine_shapes(cfg)
        self.offsets: dict[str, int] = {}
        off = 0
        for name, shape in self.shapes.items():
            self.offsets[name] = off
            n = math.prod(shape)
            off += (n + ALIGN - 1) // ALIGN * ALIGN
            if name == self.kernel_names[-1]:
                self.kernel_end = off
        self.total = off
        # tensors whose gradients are accumulated with atomics / +=, i.e. must be zeroed every step:
        # everything after the embedding table plus the table itself is handled separately (LM head stores first).
        self.small_begin = self.offsets["img.patch_b"]

    def view(self, flat: torch.Tensor, name: str) -> torch.Tensor:
        shape = self.shapes[name]
        o = self.offsets[name]
        return flat[o : o + math.prod(shape)].view(shape)


def reference_to_engine(cfg: LAPConfig, ref: dict[str, torch.Tensor]) -> dict[str, torch.Tensor]:
    """Permute/stack the reference tree into engine tensors (CPU)."""
    g, e, s = cfg.gemma, cfg.expert, cfg.siglip
    W, Ls = s.width, s.depth
    img = "PaliGemma/img/"
    blk = img + "Transformer/encoderblock/"
    mha = blk + "MultiHeadDotProductAttention_0/"
    out: dict[str, torch.Tensor] = {}
    out["img.patch_w"] = ref[img + "embedding/kernel"].reshape(-1, W).t()
    out["img.patch_b"] = ref[img + "embedding/bias"]
    out["img.pos"] = ref[img + "pos_embedding"][0]
    qkv_w = [ref[mha + f"{n}/kernel"].reshape(Ls, W, W).transpose(1, 2) for n in ("query", "key", "value")]
    out["img.qkv_w"] = torch.cat(qkv_w, dim=1)
    out["img.qkv_b"] = torch.cat([ref[mha + f"{n}/bias"].reshape(Ls, W) for n in ("query", "key", "value")], dim=1)
    out["img.out_w"] = ref[mha + "out/kernel"].reshape(Ls, W, W).transpose(1, 2)
    out["img.out_b"] = ref[mha + "out/bias"]
    out["img.fc1_w"] = ref[blk + "MlpBlock_0/Dense_0/kernel"].transpose(1, 2)
    out["img.fc1_b"] = ref[blk + "MlpBlock_0/Dense_0/bias"]
    out["img.fc2_w"] = ref[blk + "MlpBlock_0/Dense_1/kernel"].transpose(1, 2)
    out["img.fc2_b"] = ref[blk + "MlpBlock_0/Dense_1/bias"]
    out["img.ln0_s"], out["img.ln0_b"] = ref[blk + "LayerNorm_0/scale"], ref[blk + "LayerNorm_0/bias"]
    out["img.ln1_s"], out["img.ln1_b"] = ref[blk + "LayerNorm_1/scale"], ref[blk + "LayerNorm_1/bias"]
    out["img.enc_s"] = ref[img + "Transformer/encoder_norm/scale"]
    out["img.enc_b"] = ref[img + "Transformer/encoder_norm/bias"]
    out["img.head_w"] = ref[img + "head/kernel"].t()
    out["img.head_b"] = ref[img + "head/bias"]
    llm = "PaliGemma/llm/"
    lay = llm + "layers/"
    out["g.embed"] = ref[llm + "embedder/input_embedding"]
    for pre, sfx, c in (("g.", "", g), ("e.", "_1", e)):
        L, D, hd = c.depth, c.width, c.head_dim
        q = ref[lay + f"attn/q_einsum{sfx}/w"]  # [L,N,D,H] -> rows (n,h), cols d
        q = q.permute(0, 1, 3, 2).reshape(L, c.num_heads * hd, D)
        kv = ref[lay + f"attn/kv_einsum{sfx}/w"]  # [L,2,1,D,H]
        k = kv[:, 0, 0].transpose(1, 2)
        v = kv[:, 1, 0].transpose(1, 2)
        out[pre + "qkv_w"] = torch.cat([q, k, v], dim=1)
        o = ref[lay + f"attn/attn_vec_einsum{sfx}/w"]  # [L,N,H,D] -> [L, D, N*H]
        out[pre + "o_w"] = o.reshape(L, c.num_heads * hd, D).transpose(1, 2)
        gu = ref[lay + f"mlp{sfx}/gating_einsum"]  # [L,2,D,F] -> [L, 2F, D]
        out[pre + "gu_w"] = gu.transpose(2, 3).reshape(L, 2 * c.mlp_dim, D)
        out[pre + "down_w"] = ref[lay + f"mlp{sfx}/linear"].transpose(1, 2)
    out["g.attn_norm_s"] = ref[lay + "pre_attention_norm/scale"]
    out["g.ffn_norm_s"] = ref[lay + "pre_ffw_norm/scale"]
    out["g.final_norm_s"] = ref[llm + "final_norm/scale"]
    # adaRMS modulation Dense layers stacked: index 2l = pre_attention_norm_1[l], 2l+1 = pre_ffw_norm_1[l], last = final
    ka, kf = ref[lay + "pre_attention_norm_1/Dense_0/kernel"], ref[lay + "pre_ffw_norm_1/Dense_0/kernel"]
    ba, bf = ref[lay + "pre_attention_norm_1/Dense_0/bias"], ref[lay + "pre_ffw_norm_1/Dense_0/bias"]
    L = e.depth
    mw = torch.stack([ka, kf], dim=1).reshape(2 * L, e.width, 3 * e.width)
    mw = torch.cat([mw, ref[llm + "final_norm_1/Dense_0/kernel"][None]], dim=0)
    out["e.mod_w"] = mw.transpose(1, 2)
    mb = torch.stack([ba, bf], dim=1).reshape(2 * L, 3 * e.width)
    out["e.mod_b"] = torch.cat([mb, ref[llm + "final_norm_1/Dense_0/bias"][None]], dim=0)
    out["action_in_w"] = ref["action_in_proj/kernel"].t()
    out["action_in_b"] = ref["action_in_proj/bias"]
    out["time_in_w"] = ref["time_mlp_in/kernel"].t()
    out["time_in_b"] = ref["time_mlp_in/bias"]
    out["time_out_w"] = ref["time_mlp_out/kernel"].t()
    out["time_out_b"] = ref["time_mlp_out/bias"]
    out["action_out_w"] = ref["action_out_proj/kernel"].t()
    out["action_out_b"] = ref["action_out_proj/bias"]
    return {k: v.contiguous() for k, v in out.items()}


def engine_to_reference(cfg: LAPConfig, eng: dict[str, torch.Tensor]) -> dict[str, torch.Tensor]:
    """Inverse of reference_to_engine (used for checkpoints and for comparing gradients with the oracle)."""
    g, e, s = cfg.gemma, cfg.expert, cfg.siglip
    W, Ls, nh, hd = s.width, s.depth, s.num_heads, s.head_dim
    ps = s.patch_size
    img = "PaliGemma/img/"
    blk = img + "Transformer/encoderblock/"
    mha = blk + "MultiHeadDotProductAttention_0/"
    ref: dict[str, torch.Tensor] = {}
    ref[img + "embedding/kernel"] = eng["img.patch_w"].t().reshape(ps, ps, 3, W)
    ref[img + "embedding/bias"] = eng["img.patch_b"]
    ref[img + "pos_embedding"] = eng["img.pos"][None]
    for i, n in enumerate(("query", "key", "value")):
        ref[mha + f"{n}/kernel"] = eng["img.qkv_w"][:, i * W : (i + 1) * W].transpose(1, 2).reshape(Ls, W, nh, hd)
        ref[mha + f"{n}/bias"] = eng["img.qkv_b"][:, i * W : (i + 1) * W].reshape(Ls, nh, hd)
    ref[mha + "out/kernel"] = eng["img.out_w"].transpose(1, 2).reshape(Ls, nh, hd, W)
    ref[mha + "out/bias"] = eng["img.out_b"]
    ref[blk + "MlpBlock_0/Dense_0/kernel"] = eng["img.fc1_w"].transpose(1, 2)
    ref[blk + "MlpBlock_0/Dense_0/bias"] = eng["img.fc1_b"]
    ref[blk + "MlpBlock_0/Dense_1/kernel"] = eng["img.fc2_w"].transpose(1, 2)
    ref[blk + "MlpBlock_0/Dense_1/bias"] = eng["img.fc2_b"]
    ref[blk + "LayerNorm_0/scale"], ref[blk + "LayerNorm_0/bias"] = eng["img.ln0_s"], eng["img.ln0_b"]
    ref[blk + "LayerNorm_1/scale"], ref[blk + "LayerNorm_1/bias"] = eng["img.ln1_s"], eng["img.ln1_b"]
    ref[img + "Transformer/encoder_norm/scale"] = eng["img.enc_s"]
    ref[img + "Transformer/encoder_norm/bias"] = eng["img.enc_b"]
    ref[img + "head/kernel"] = eng["img.head_w"].t()
    ref[img + "head/bias"] = eng["img.head_b"]
    llm = "PaliGemma/llm/"
    lay = llm + "layers/"
    ref[llm + "embedder/input_embedding"] = eng["g.embed"]
    for pre, sfx, c in (("g.", "", g), ("e.", "_1", e)):
        L, D, H = c.depth, c.width, c.head_dim
        nq = c.num_heads * H
        qkv = eng[pre + "qkv_w"]
        ref[lay + f"attn/q_einsum{sfx}/w"] = qkv[:, :nq].reshape(L, c.num_heads, H, D).permute(0, 1, 3, 2)
        k = qkv[:, nq : nq + H].transpose(1, 2)
        v = qkv[:, nq + H :].transpose(1, 2)
        ref[lay + f"attn/kv_einsum{sfx}/w"] = torch.stack([k, v], dim=1)[:, :, None]
        ref[lay + f"attn/attn_vec_einsum{sfx}/w"] = eng[pre + "o_w"].transpose(1, 2).reshape(L, c.num_heads, H, D)
        ref[lay + f"mlp{sfx}/gating_einsum"] = eng[pre + "gu_w"].reshape(L, 2, c.mlp_dim, D).transpose(2, 3)
        ref[lay + f"mlp{sfx}/linear"] = eng[pre + "down_w"].transpose(1, 2)
    ref[lay + "pre_attention_norm/scale"] = eng["g.attn_norm_s"]
    ref[lay + "pre_ffw_norm/scale"] = eng["g.ffn_norm_s"]
    ref[llm + "final_norm/scale"] = eng["g.final_norm_s"]
    L = e.depth
    mw = eng["e.mod_w"].transpose(1, 2)  # [37, D1, 3D1]
    ref[lay + "pre_attention_norm_1/Dense_0/kernel"] = mw[: 2 * L : 2]
    ref[lay + "pre_ffw_norm_1/Dense_0/kernel"] = mw[1 : 2 * L : 2]
    ref[llm + "final_norm_1/Dense_0/kernel"] = mw[2 * L]
    mb = eng["e.mod_b"]
    ref[lay + "pre_attention_norm_1/Dense_0/bias"] = mb[: 2 * L : 2]
    ref[lay + "pre_ffw_norm_1/Dense_0/bias"] = mb[1 : 2 * L : 2]
    ref[llm + "final_norm_1/Dense_0/bias"] = mb[2 * L]
    ref["action_in_proj/kernel"] = eng["action_in_w"].t()
    ref["action_in_proj/bias"] = eng["action_in_b"]
    ref["time_mlp_in/kernel"] = eng["time_in_w"].t()
    ref["time_mlp_in/bias"] = eng["time_in_b"]
    ref["time_mlp_out/kernel"] = eng["time_out_w"].t()
    ref["time_mlp_out/bias"] = eng["time_out_b"]
    ref["action_out_proj/kernel"] = eng["action_out_w"].t()
    ref["action_out_proj/bias"] = eng["action_out_b"]
    return {k: v.contiguous() for k, v in ref.items()}


def flat_from_engine(layout: FlatLayout, eng: dict[str, torch.Tensor], device=None) -> torch.Tensor:
    flat = torch.zeros(layout.total, dtype=torch.float32, device=device)
    for name in layout.shapes:
        layout.view(flat, name).copy_(eng[name].to(flat.device))
    return flat


def engine_from_flat(layout: FlatLayout, flat: torch.Tensor) -> dict[str, torch.Tensor]:
    return {name: layout.view(flat, name) for name in layout.shapes}


def to_nested(flat_tree: dict[str, torch.Tensor]) -> dict:
    """'/'-joined keys -> nested dicts (the shape orbax / nnx `to_pure_dict` uses)."""
    root: dict = {}
    for k, v in flat_tree.items():
        parts = k.split("/")
        d = root
        for p in parts[:-1]:
            d = d.setdefault(p, {})
        d[parts[-1]] = v
    return root


def from_nested(tree: dict, prefix: str = "") -> dict[str, torch.Tensor]:
    """Nested dicts -> '/'-joined keys; strips the trailing `value` level some checkpoints carry (model.py:286-332)."""
    out: dict[str, torch.Tensor] = {}
    for k, v in tree.items():
        key = f"{prefix}/{k}" if prefix else str(k)
        if isinstance(v, dict):
            if set(v.keys()) == {"value"}:
                out[key] = v["value"]
            else:
                out.update(from_nested(v, key))
        else:
            out[key] = v
    return out
