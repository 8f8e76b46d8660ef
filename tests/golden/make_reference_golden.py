"""Generate tests/golden/reference_pi05_*.npz by RUNNING THE REFERENCE'S OWN CODE in this container.

The JAX/Flax implementation of the hot path cannot run here (no jax/flax wheels), but the reference
vendors a second, PyTorch statement of the same π0.5 arithmetic that LAP shares line-for-line when
`stop_action_to_vlm_grad=False`:

  third_party/openpi/src/openpi/models_pytorch/pi0_pytorch.py          (PI0Pytorch: embed_prefix / embed_suffix / forward /
                                                                         sample_actions / denoise_step)
  third_party/openpi/src/openpi/models_pytorch/gemma_pytorch.py        (PaliGemmaWithExpertModel: joint two-expert layers)
  third_party/openpi/src/openpi/models_pytorch/transformers_replace/** (patched HF Gemma adaRMS + gated residual, SigLIP,
                                                                         PaliGemma)
  third_party/openpi/examples/convert_jax_model_to_pytorch.py          (the JAX param tree -> PyTorch state-dict mapping)

This script imports those files UNMODIFIED from /root/reference, feeds them a seeded parameter tree in the reference's JAX
layout (mapped by the reference's own converter functions), and records inputs/outputs.  tests/test_reference_golden.py
then checks oracle/lap_oracle.py (configured as π0.5: no lang-action tokens, three cameras, action_dim 32, fp32) against
these outputs — pinning oracle rows a9, a10, a11, a12(encode), a13(make_attn_mask), a14–a18, a20, a21 of SURVEY §8(a).
Not pinned by this (LAP-specific, no second statement anywhere in the reference): the lang-action mask rows, the language
CE loss and the loss weighting.

Environment shims (none touches reference arithmetic; all are listed so a reader can audit them):
  S1  the reference's patched transformers files are overlaid onto the installed transformers 5.5 in sys.modules (the
      reference's install step is `cp -r transformers_replace/* site-packages/transformers/`); transformers.utils.LossKwargs and
      transformers.cache_utils.HybridCache, which 4.53.2 had and 5.5 dropped, are defined as empty types for the import.
  S2  `transformers.__version__` reads "4.53.2" while PI0Pytorch.__init__ runs its installed-correctly check.
  S3  `openpi.models.gemma` (Flax; needs jax) is replaced by a stub exposing only `get_config` -> the hyper-parameter record
      (width, depth, mlp_dim, num_heads, num_kv_heads, head_dim) of the small test variants; `openpi.shared.image_tools` (jax
      resize, unused at 224x224) is an empty stub.
  S4  hyper-parameters PaliGemmaWithExpertModel hard-codes for the 3B model are shrunk just before the HF modules are built:
      vocab 257152 -> VOCAB, SigLIP So400m -> (VIS_WIDTH, VIS_DEPTH, VIS_HEADS, VIS_MLP), projection_dim 2048 -> gemma width.
  S5  image augmentation is off (PI0Pytorch.forward passes train=True to its preprocessing; lap_libero sets
      enable_image_augmentation=False, src/lap/training/config.py:761) by forcing train=False in that one call.
  S7  transformers 5.5 dropped the "default" entry of ROPE_INIT_FUNCTIONS that the reference's GemmaRotaryEmbedding looks
      up (modeling_gemma.py:141); it is restated from transformers 4.53.2 (inv_freq_i = rope_theta^(-2i/head_dim), scale 1.0)
      — the one piece of third-party arithmetic in this pin, cross-checked against gemma.py:548-564 `_apply_rope`
      (max_wavelength 10_000) by the oracle agreeing with the result.
  S8  tie_word_embeddings=False on the HF configs (5.5's weight-tying bookkeeping cannot read 4.53.2's list-style
      `_tied_weights_keys`; the tied lm_head is never called on this path).
  S9  DynamicCache.__getitem__(layer) -> (keys, values), which 4.53.2 had and the reference's attention reads
      (modeling_gemma.py:309-310), is added back on top of 5.5's per-layer storage.
  S6  sample_actions is called as the plain method (the reference wraps it in torch.compile(mode="max-autotune") in
      __init__, which would need a GPU toolchain; torch.compile does not change semantics).

Run:  python tests/golden/make_reference_golden.py          (needs /root/reference; writes tests/golden/reference_pi05_*.npz)
"""
from __future__ import annotations

import ast
import dataclasses
import hashlib
import importlib
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for _p in (ROOT, HERE):
    if _p not in sys.path:
        sys.path.insert(0, _p)

REF = os.environ.get("LAP_REFERENCE", "/root/reference")
OP_SRC = os.path.join(REF, "third_party/openpi/src")
TR = os.path.join(OP_SRC, "openpi/models_pytorch/transformers_replace/models")
CONVERTER = os.path.join(REF, "third_party/openpi/examples/convert_jax_model_to_pytorch.py")

from reference_cases import grad_fingerprint  # noqa: E402
from reference_cases import (ACTION_DIM, CASES, VIS_DEPTH, VIS_HEADS, VIS_MLP, VIS_WIDTH, VOCAB, lap_config, pack_rows,  # noqa: E402
                             params_digest, seeded_inputs, seeded_reference_params)

ROW_STRIDE = 9


# ------------------------------------------------------------------------------------------------------------------
# loading the reference (shims S1-S6)
# ------------------------------------------------------------------------------------------------------------------
def _overlay(modname: str, path: str):
    spec = importlib.util.spec_from_file_location(modname, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[modname] = m
    spec.loader.exec_module(m)
    parent, _, leaf = modname.rpartition(".")
    setattr(importlib.import_module(parent), leaf, m)
    return m


def load_reference_pytorch_port():
    import transformers
    import transformers.cache_utils
    import transformers.utils
    from typing import TypedDict

    # S1
    if not hasattr(transformers.utils, "LossKwargs"):
        class LossKwargs(TypedDict, total=False):
            num_items_in_batch: int
        transformers.utils.LossKwargs = LossKwargs
    if not hasattr(transformers.cache_utils, "HybridCache"):
        class HybridCache(transformers.cache_utils.Cache):
            pass
        transformers.cache_utils.HybridCache = HybridCache
    if not hasattr(transformers.cache_utils.DynamicCache, "__getitem__"):  # S9
        def _getitem(self, layer_idx):
            layer = self.layers[layer_idx]
            return layer.keys, layer.values
        transformers.cache_utils.DynamicCache.__getitem__ = _getitem
    for mod, rel in [("transformers.models.gemma.configuration_gemma", "gemma/configuration_gemma.py"),
                     ("transformers.models.gemma.modeling_gemma", "gemma/modeling_gemma.py"),
                     ("transformers.models.siglip.check", "siglip/check.py"),
                     ("transformers.models.siglip.modeling_siglip", "siglip/modeling_siglip.py"),
                     ("transformers.models.paligemma.modeling_paligemma", "paligemma/modeling_paligemma.py")]:
        _overlay(mod, os.path.join(TR, rel))
    # the top-level names gemma_pytorch.py imports must be the overlaid classes
    from transformers.models.auto import CONFIG_MAPPING
    mg = sys.modules["transformers.models.gemma.modeling_gemma"]
    mp = sys.modules["transformers.models.paligemma.modeling_paligemma"]
    cg = sys.modules["transformers.models.gemma.configuration_gemma"]
    if not hasattr(cg.GemmaConfig, "use_bidirectional_attention"):
        cg.GemmaConfig.use_bidirectional_attention = None  # field transformers 5.5's PaliGemmaConfig reads; unused by 4.53.2 code
    if "default" not in mg.ROPE_INIT_FUNCTIONS:  # S7
        def _default_rope(config=None, device=None, seq_len=None, **_):
            # transformers 4.53.2 modeling_rope_utils._compute_default_rope_parameters (third-party, not under
            # /root/reference; restated): inv_freq_i = base^(-2i/dim), attention_factor 1.0
            base = config.rope_theta
            dim = getattr(config, "head_dim", None) or config.hidden_size // config.num_attention_heads
            inv_freq = 1.0 / (base ** (torch.arange(0, dim, 2, dtype=torch.int64).to(device=device, dtype=torch.float) / dim))
            return inv_freq, 1.0
        mg.ROPE_INIT_FUNCTIONS = dict(mg.ROPE_INIT_FUNCTIONS, default=_default_rope)
    transformers.GemmaForCausalLM = mg.GemmaForCausalLM
    transformers.PaliGemmaForConditionalGeneration = mp.PaliGemmaForConditionalGeneration

    # S3
    from lap_b200.config import get_gemma_config

    if OP_SRC not in sys.path:
        sys.path.insert(0, OP_SRC)
    for pkg in ("openpi", "openpi.models", "openpi.shared", "openpi.models_pytorch"):
        if pkg not in sys.modules:
            m = types.ModuleType(pkg)
            m.__path__ = [os.path.join(OP_SRC, *pkg.split("."))]
            sys.modules[pkg] = m
    g = types.ModuleType("openpi.models.gemma")
    g.get_config = get_gemma_config
    sys.modules["openpi.models.gemma"] = g
    sys.modules["openpi.models"].gemma = g
    it = types.ModuleType("openpi.shared.image_tools")
    sys.modules["openpi.shared.image_tools"] = it
    sys.modules["openpi.shared"].image_tools = it

    gp = importlib.import_module("openpi.models_pytorch.gemma_pytorch")
    pp = importlib.import_module("openpi.models_pytorch.pi0_pytorch")

    # the reference's patched GemmaConfig must be the one CONFIG_MAPPING["gemma"] builds (use_adarms / adarms_cond_dim)
    class _Mapping(dict):
        pass
    real_pg_cfg = CONFIG_MAPPING["paligemma"]
    mapping = _Mapping(paligemma=real_pg_cfg, gemma=cg.GemmaConfig)
    gp.CONFIG_MAPPING = mapping

    # S4
    real_pg = gp.PaliGemmaForConditionalGeneration
    real_gemma = gp.GemmaForCausalLM

    def small_paligemma(config):
        width = config.text_config.hidden_size
        # the text tower must be built from the reference's patched GemmaConfig fields
        config.text_config.vocab_size = VOCAB
        config._vocab_size = VOCAB
        config.vocab_size = VOCAB
        config.image_token_index = VOCAB
        config.text_config.pad_token_id = 0
        config.pad_token_id = 0
        config.vision_config.hidden_size = VIS_WIDTH
        config.vision_config.num_hidden_layers = VIS_DEPTH
        config.vision_config.num_attention_heads = VIS_HEADS
        config.vision_config.intermediate_size = VIS_MLP
        config.vision_config.projection_dim = width
        config.vision_config.vision_use_head = False
        config.projection_dim = width
        # S8: lm_head tying is bookkeeping transformers 5.5 does differently (dict vs list); the LM head is not on this path
        config.tie_word_embeddings = False
        config.text_config.tie_word_embeddings = False
        return real_pg(config=config)

    def small_gemma(config):
        config.vocab_size = VOCAB
        config.pad_token_id = 0
        config.tie_word_embeddings = False  # S8
        return real_gemma(config=config)

    gp.PaliGemmaForConditionalGeneration = small_paligemma
    gp.GemmaForCausalLM = small_gemma

    # S5
    real_pre = pp._preprocessing.preprocess_observation_pytorch
    pp._preprocessing.preprocess_observation_pytorch = lambda obs, *, train=False, **kw: real_pre(obs, train=False, **kw)
    return pp


def load_reference_converter():
    """The two mapping functions of examples/convert_jax_model_to_pytorch.py, executed from the reference's source
    (the module itself imports jax/orbax at the top, so the function definitions are compiled out of its AST)."""
    src = open(CONVERTER).read()
    tree = ast.parse(src)
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef)
            and n.name in ("slice_paligemma_state_dict", "slice_gemma_state_dict")]
    assert len(keep) == 2
    ns = {"np": np, "torch": torch}
    exec(compile(ast.Module(body=keep, type_ignores=[]), CONVERTER, "exec"), ns)
    return ns["slice_paligemma_state_dict"], ns["slice_gemma_state_dict"]


@dataclasses.dataclass
class RefModelConfig:
    """What PI0Pytorch reads off its config (pi0_pytorch.py:84-100, 330, 378)."""
    pi05: bool
    paligemma_variant: str
    action_expert_variant: str
    dtype: str
    action_horizon: int
    action_dim: int


class RefObservation:
    """Attribute bag with the fields preprocessing_pytorch.py:19-173 reads."""

    def __init__(self, cfg, inputs):
        t = torch.from_numpy
        self.images = {k: t(inputs["image/" + k]).permute(0, 3, 1, 2).contiguous() for k in cfg.image_keys}
        self.image_masks = {k: t(inputs["image_mask/" + k]) for k in cfg.image_keys}
        self.state = t(inputs["state"])
        self.tokenized_prompt = t(inputs["tokenized_prompt"]).long()
        self.tokenized_prompt_mask = t(inputs["tokenized_prompt_mask"])
        self.token_ar_mask = None
        self.token_loss_mask = None


def build_reference_model(pp, cfg, params: dict[str, np.ndarray]):
    import transformers

    rc = RefModelConfig(pi05=True, paligemma_variant=cfg.paligemma_variant, action_expert_variant=cfg.action_expert_variant,
                        dtype="float32", action_horizon=cfg.action_horizon, action_dim=cfg.action_dim)
    real_version = transformers.__version__
    transformers.__version__ = "4.53.2"  # S2
    try:
        model = pp.PI0Pytorch(rc)
    finally:
        transformers.__version__ = real_version
    slice_paligemma, slice_gemma = load_reference_converter()
    pg = {k[len("PaliGemma/"):]: v.copy() for k, v in params.items() if k.startswith("PaliGemma/")}
    hf = model.paligemma_with_expert.paligemma.config
    paligemma_params, expert_params = slice_paligemma(pg, hf)
    expert_cfg = types.SimpleNamespace(**dataclasses.asdict(cfg.expert))  # the converter adds attributes to it (:274-281)
    gemma_params = slice_gemma(expert_params, expert_cfg, num_expert=1, checkpoint_dir="pi05", pi05=True)
    proj = {}
    for key in ("action_in_proj", "action_out_proj", "time_mlp_in", "time_mlp_out"):  # converter :437-470
        proj[key + ".weight"] = torch.from_numpy(params[key + "/kernel"]).T
        proj[key + ".bias"] = torch.from_numpy(params[key + "/bias"])
    allp = {**paligemma_params, **gemma_params, **proj}
    missing, unexpected = model.load_state_dict(allp, strict=False)
    # the only tensors the converter legitimately leaves untouched: tied / unused heads and the unconditioned RMSNorm
    # weights of adaRMS layers
    bad = [k for k in missing if not (k.endswith("lm_head.weight") or "vision_model.head" in k)]
    assert not bad, f"reference converter left parameters unset: {bad}"
    assert not unexpected, f"unexpected keys: {unexpected}"
    model.eval()
    return model


def torch_index_of_jax_layout(cfg, params: dict[str, np.ndarray]) -> dict[str, tuple[str, np.ndarray]]:
    """Which JAX-tree element every PyTorch-port parameter element comes from, WITHOUT hand-writing the inverse of the
    reference's converter: the converter (pure transposes / reshapes / slices) is applied to a tree of element ids."""
    slice_paligemma, slice_gemma = load_reference_converter()
    ids, offs, off = {}, {}, 0
    for k, v in params.items():
        ids[k] = (off + np.arange(v.size, dtype=np.float64)).reshape(v.shape)
        offs[k] = off
        off += v.size
    pg = {k[len("PaliGemma/"):]: v.copy() for k, v in ids.items() if k.startswith("PaliGemma/")}
    hf = types.SimpleNamespace(
        vision_config=types.SimpleNamespace(hidden_size=VIS_WIDTH, num_hidden_layers=VIS_DEPTH),
        text_config=types.SimpleNamespace(hidden_size=cfg.gemma.width, num_hidden_layers=cfg.gemma.depth,
                                          num_attention_heads=cfg.gemma.num_heads, head_dim=cfg.gemma.head_dim))
    pal, exp = slice_paligemma(pg, hf)
    gem = slice_gemma(exp, types.SimpleNamespace(**dataclasses.asdict(cfg.expert)), num_expert=1, checkpoint_dir="pi05",
                      pi05=True)
    out = {**{k: v.numpy() for k, v in pal.items()}, **{k: v.numpy() for k, v in gem.items()}}
    for key in ("action_in_proj", "action_out_proj", "time_mlp_in", "time_mlp_out"):
        out[key + ".weight"] = ids[key + "/kernel"].T
        out[key + ".bias"] = ids[key + "/bias"]
    return out, offs, off


def reference_grad_summary(pp, cfg, params, model, obs, actions, noise, time) -> dict[str, np.ndarray]:
    for p_ in model.parameters():
        p_.grad = None
    with torch.enable_grad():
        loss = model.forward(obs, actions, noise=noise, time=time).mean()
        loss.backward()
    idx, offs, total = torch_index_of_jax_layout(cfg, params)
    flat = np.zeros(total, dtype=np.float64)
    named = dict(model.named_parameters())
    for k, id_arr in idx.items():
        if k in named and named[k].grad is not None:
            np.add.at(flat, id_arr.astype(np.int64).reshape(-1), named[k].grad.detach().double().numpy().reshape(-1))
    out = {"grad_loss": np.float32(loss.item())}
    for k, v in params.items():
        g = flat[offs[k]: offs[k] + v.size]
        out["grad/" + k] = grad_fingerprint(g)
    return out


def run_case(pp, case: str) -> dict[str, np.ndarray]:
    pv, ev, batch, horizon, L, seed = CASES[case]
    cfg = lap_config(case)
    params = seeded_reference_params(cfg, seed)
    inputs = seeded_inputs(cfg, batch, seed)
    model = build_reference_model(pp, cfg, params)
    obs = RefObservation(cfg, inputs)
    t = torch.from_numpy
    out = {"params_sha256": np.frombuffer(params_digest(params).encode(), dtype=np.uint8)}
    with torch.no_grad():
        # a11: SigLIP tower + head, per camera
        img = obs.images["base_0_rgb"]
        out["siglip_tokens"] = model.paligemma_with_expert.embed_image(img).numpy()
        # a10/a12: prefix tokens + masks
        images, img_masks, lang_tokens, lang_masks, state = model._preprocess_observation(obs, train=False)
        pe, pm, pa = model.embed_prefix(images, img_masks, lang_tokens, lang_masks)
        out["prefix_tokens"], out["prefix_pad_mask"], out["prefix_ar_mask"] = pe.numpy(), pm.numpy(), pa.numpy()
        # a8/a9: suffix tokens + adaRMS condition
        time = t(inputs["time"])
        noise = t(inputs["noise"])
        actions = t(inputs["actions"])
        te = time[:, None, None]
        x_t = te * noise + (1 - te) * actions
        se, sm, sa, cond = model.embed_suffix(state, x_t, time)
        out["suffix_tokens"], out["adarms_cond"] = se.numpy(), cond.numpy()
        # a13: mask + positions exactly as PI0Pytorch.forward builds them
        pad = torch.cat([pm, sm], 1)
        att = torch.cat([pa, sa.to(torch.bool)], 1)
        out["attn_mask"] = pp.make_att_2d_masks(pad, att).numpy()
        out["positions"] = (torch.cumsum(pad, 1) - 1).numpy()
        # a14-a18: joint two-expert transformer
        m4 = model._prepare_attention_masks_4d(pp.make_att_2d_masks(pad, att))
        (pre_out, suf_out), _ = model.paligemma_with_expert.forward(
            attention_mask=m4, position_ids=torch.cumsum(pad, 1) - 1, past_key_values=None, inputs_embeds=[pe, se],
            use_cache=False, adarms_cond=[None, cond])
        out["prefix_out"], out["suffix_out"] = pre_out.numpy(), suf_out.numpy()
        # a5/a20: the training forward (per-element squared error of the vector field)
        out["mse"] = model.forward(obs, actions, noise=noise, time=time).numpy()
        # a21: inference
        out["sampled_actions"] = pp.PI0Pytorch.sample_actions.__wrapped__(model, torch.device("cpu"), obs, noise=noise,
                                                                           num_steps=10).numpy()  # S6
        out["sampled_actions_3"] = pp.PI0Pytorch.sample_actions.__wrapped__(model, torch.device("cpu"), obs, noise=noise,
                                                                             num_steps=3).numpy()
    # backward: d mean(mse) / d params through the reference's autograd, scattered back to the JAX layout and summarised
    out.update(reference_grad_summary(pp, cfg, params, model, obs, actions, noise, time))
    # module-level helpers
    out["posemb_t"] = np.linspace(0.001, 1.0, 7, dtype=np.float32)
    out["posemb"] = pp.create_sinusoidal_pos_embedding(t(out["posemb_t"]), 32, 4e-3, 4.0, device=torch.device("cpu")).numpy()
    return out


def main():
    torch.manual_seed(0)
    torch.set_num_threads(max(1, (os.cpu_count() or 2) // 2))
    pp = load_reference_pytorch_port()
    for case in CASES:
        out = run_case(pp, case)
        path = os.path.join(HERE, f"reference_pi05_{case}.npz")
        out = {k: (v.astype(np.float32) if v.dtype == np.float64 else v) for k, v in out.items()}
        for k in ("siglip_tokens", "prefix_tokens", "prefix_out"):  # keep fixtures small: strided token rows
            out[k] = pack_rows(out[k], ROW_STRIDE)
        out["row_stride"] = np.int64(ROW_STRIDE)
        out["attn_mask"] = np.packbits(out["attn_mask"], axis=-1)
        np.savez_compressed(path, **out)
        print(case, {k: v.shape for k, v in out.items()}, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
