"""Generate tests/golden/reference_lap_*.npz: the LAP-specific half of the hot path, from the reference's own source.

make_reference_golden.py pins what LAP shares with π0.5 by running the reference's PyTorch port.  What is LAP-only and has
no PyTorch statement — the lang-action prefix-LM mask and action-rows-skip-langact mask (`lap.py:303-377`), `embed_prefix`
with `tokenized_langact_mask` as AR mask (`lap.py:118-170`), `prepare_suffix` (`:185-207`), the language CE loss
(`:209-289`), the action loss (`:291-301`), the loss weighting / normalisation of `compute_loss` (`:380-602`) and
`sample_actions` (`:605-675`, incl. its quirk of letting action rows see lang-action keys) and the greedy autoregressive
`sample_tokens` (`:678-766`, with `pi0_fast.left_to_right_align / put_along_last_axis`) — is JAX code.  Its *leaf*
modules (SigLIP, the two-expert Gemma stack, the nnx.Linear projections) are exactly the ones the PyTorch port restates, and
everything between the leaves is plain `jax.numpy` array algebra.  So this script executes the reference's method bodies
as they stand in /root/reference (compiled from the source files' AST, annotations and decorators dropped because they
need jaxtyping/flax at definition time), with

  * `jnp`   -> numpy (fp32 in, fp32 out; `.at[idx].set(v)` provided by an ndarray subclass),
  * `jax.random.normal/beta` -> the explicit `noise` / `time` the test also feeds the oracle (JAX's threefry stream cannot be
    reproduced), `jax.random.split` -> dummy keys, `jax.lax.while_loop` -> a Python while loop, `jax.lax.cond` -> if/else,
    `jax.vmap` -> a per-example loop,
  * `jax.nn.one_hot / log_softmax`, `nnx.swish`, `einops` -> their textbook definitions (third-party: jax 0.5.3 / flax 0.10.2,
    not under /root/reference),
  * `preprocess_observation` -> identity (lap_libero: `enable_image_augmentation=False`, images already 224x224),
  * `self.PaliGemma.img`, `self.PaliGemma.llm` (call / method="embed" / kv-cache) and the four nnx.Linear projections -> the
    reference's PyTorch port modules, fp32, loaded through the reference's own JAX->PyTorch converter
    (make_reference_golden.build_reference_model);  `llm(..., method="decode")` -> `x @ E.T` (`gemma.py:153-154`).

tests/test_reference_golden.py checks oracle/lap_oracle.py (and, on the GPU, the engine) against what this records.
Run:  python tests/golden/make_reference_lap_golden.py      (needs /root/reference)
"""
from __future__ import annotations

import ast
import logging
import math
import os
import sys
import types

import einops
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_reference_golden as G  # noqa: E402
from reference_cases import LAP_CASES, lap_case_config, lap_case_inputs, pack_rows, params_digest, seeded_reference_params  # noqa: E402

LAP_PY = os.path.join(G.REF, "src/lap/models/lap.py")
PI0_PY = os.path.join(G.OP_SRC, "openpi/models/pi0.py")
METRICS_PY = os.path.join(G.REF, "src/lap/models/model_utils/metrics.py")


# ------------------------------------------------------------------------------------------------------------------
# numpy standing in for jax.numpy / jax
# ------------------------------------------------------------------------------------------------------------------
class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        arr = self.arr

        class _Setter:
            def set(self, value):
                out = arr.copy()
                out[idx] = value
                return out.view(AtArray)
        return _Setter()


class AtArray(np.ndarray):
    @property
    def at(self):
        return _At(self)


def _wrap(x):
    return np.asarray(x).view(AtArray)


class _Jnp:
    """jax.numpy names used by the executed reference code, forwarded to numpy."""
    pi = np.pi
    float32 = np.float32
    int32 = np.int32
    bool_ = np.bool_

    def __getattr__(self, name):
        return getattr(np, name)

    @staticmethod
    def zeros(shape, dtype=np.float32):
        return _wrap(np.zeros(shape, dtype=dtype))

    @staticmethod
    def ones(shape, dtype=np.float32):
        return _wrap(np.ones(shape, dtype=dtype))

    @staticmethod
    def linspace(a, b, n):
        return np.linspace(a, b, n, dtype=np.float32)  # jnp default dtype is float32

    @staticmethod
    def einsum(eq, *ops, precision=None):
        return np.einsum(eq, *[np.asarray(o, dtype=np.float32) for o in ops])

    @staticmethod
    def array(x, dtype=None):
        return np.array(x, dtype=dtype)

    @staticmethod
    def asarray(x, dtype=None):
        return np.asarray(x, dtype=dtype)

    @staticmethod
    def clip(x, lo=None, hi=None):
        return np.clip(x, lo, hi)


def make_jax_shim(noise, time):
    jax = types.SimpleNamespace()
    jax.random = types.SimpleNamespace(
        split=lambda rng, n=2: [None] * n,
        normal=lambda rng, shape: np.asarray(noise, dtype=np.float32).reshape(shape),
        # prepare_suffix computes beta * 0.999 + 0.001; hand back the beta draw that yields exactly `time`
        beta=lambda rng, a, b, shape: _BetaDraw(np.asarray(time, dtype=np.float32).reshape(shape)),
    )

    def one_hot(x, n, dtype=np.float32):
        return (np.asarray(x)[..., None] == np.arange(n)).astype(dtype)

    def log_softmax(x, axis=-1):
        x = np.asarray(x, dtype=np.float32)
        s = x - x.max(axis=axis, keepdims=True)
        return s - np.log(np.exp(s).sum(axis=axis, keepdims=True))

    jax.nn = types.SimpleNamespace(one_hot=one_hot, log_softmax=log_softmax)

    def while_loop(cond, body, carry):
        while cond(carry):
            carry = body(carry)
        return carry

    def cond(pred, true_fn, false_fn, operand=None):
        return true_fn(operand) if pred else false_fn(operand)

    def vmap(f):  # per-example Python loop (pi0_fast.left_to_right_align is written for ONE example and vmapped)
        def g(*args):
            outs = [f(*[a[i] for a in args]) for i in range(len(args[0]))]
            return tuple(np.stack([o[k] for o in outs]) for k in range(len(outs[0])))
        return g

    jax.lax = types.SimpleNamespace(while_loop=while_loop, cond=cond, Precision=types.SimpleNamespace(HIGHEST=None))
    jax.vmap = vmap
    return jax


class _BetaDraw:
    """`draw * 0.999 + 0.001` must give back exactly the fp32 `time` the oracle is fed."""

    def __init__(self, time):
        self.time = time

    def __mul__(self, k):
        assert k == 0.999
        return self

    def __add__(self, k):
        assert k == 0.001
        return self.time


# ------------------------------------------------------------------------------------------------------------------
# reference source -> callables
# ------------------------------------------------------------------------------------------------------------------
def _strip(fn: ast.FunctionDef) -> ast.FunctionDef:
    fn.decorator_list = []
    fn.returns = None
    for a in fn.args.args + fn.args.kwonlyargs + fn.args.posonlyargs:
        a.annotation = None
    for sub in ast.walk(fn):
        if isinstance(sub, ast.FunctionDef) and sub is not fn:
            _strip(sub)
    return fn


def functions_from(path: str, names: set[str], cls: str | None = None) -> list[ast.FunctionDef]:
    tree = ast.parse(open(path).read())
    body = tree.body
    if cls is not None:
        body = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls).body
    got = [_strip(n) for n in body if isinstance(n, ast.FunctionDef) and n.name in names]
    assert {g.name for g in got} == names, (path, names - {g.name for g in got})
    return got


def exec_functions(path, fns, ns):
    exec(compile(ast.Module(body=fns, type_ignores=[]), path, "exec"), ns)
    return ns


LAP_METHODS = {"_configure_shared_training_attributes", "embed_prefix", "_embed_prefix_for_loss", "prepare_suffix",
               "_compute_language_loss", "_compute_action_loss", "_build_prefix_action_mask",
               "_build_combined_attention_mask", "_build_combined_positions", "compute_loss", "sample_actions", "sample_tokens"}
PI0_FAST_PY = os.path.join(G.OP_SRC, "openpi/models/pi0_fast.py")


class _Leaves:
    """The reference's PyTorch-port modules behind the call signatures lap.py uses."""

    def __init__(self, model, E):
        self.m, self.E, self.calls, self.decoded = model, E, [], []

    def img(self, image, train=False):
        x = torch.from_numpy(np.ascontiguousarray(image)).permute(0, 3, 1, 2).contiguous()
        with torch.no_grad():
            return self.m.paligemma_with_expert.embed_image(x).numpy(), None

    def llm(self, embedded=None, *, method=None, positions=None, mask=None, adarms_cond=None, kv_cache=None):
        if method == "embed":  # gemma.py:148-151 via Module.embed :446-448
            with torch.no_grad():
                e = self.m.paligemma_with_expert.embed_language_tokens(torch.from_numpy(np.asarray(embedded)).long())
                return (e * math.sqrt(e.shape[-1])).numpy()
        if method == "decode":  # gemma.py:153-154
            logits = np.asarray(embedded, dtype=np.float32) @ self.E.T
            self.decoded.append(logits)
            return logits
        assert method is None
        t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a))
        mask = np.asarray(mask, dtype=bool)
        if embedded[1] is None:
            # sample_tokens (lap.py:702-711, 744-750) addresses a cache PRE-ALLOCATED to prefill + max_decoding_steps slots and
            # masks the slots that are not written yet; the port's cache grows one slot per step instead, so only the columns
            # of the slots that exist are passed on (the dropped columns are all False in the prefill mask and select
            # not-yet-written slots in the decode mask - asserted)
            have = (0 if kv_cache is None else kv_cache.get_seq_length()) + np.asarray(embedded[0]).shape[1]
            if kv_cache is None:
                assert not mask[..., have:].any()
            mask = mask[..., :have]
        m4 = self.m._prepare_attention_masks_4d(t(mask))
        with torch.no_grad():
            (p, s), cache = self.m.paligemma_with_expert.forward(
                attention_mask=m4, position_ids=t(np.asarray(positions)).long(), past_key_values=kv_cache,
                inputs_embeds=[t(embedded[0]), t(embedded[1])], use_cache=(embedded[1] is None),
                adarms_cond=[None, t(adarms_cond[1])])
        self.calls.append(dict(mask=mask, positions=np.asarray(positions)))
        n = lambda a: None if a is None else a.numpy()
        return [n(p), n(s)], cache


def _linear(mod):
    def f(x):
        with torch.no_grad():
            return mod(torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))).numpy()
    f.out_features = mod.out_features
    return f


def build_reference_lap(pp, cfg, params, noise, time):
    model = G.build_reference_model(pp, cfg, params)
    jnp, jax = _Jnp(), make_jax_shim(noise, time)
    swish = lambda x: x / (1.0 + np.exp(-x))  # flax nnx.swish = x * sigmoid(x)
    pi0_ns = exec_functions(PI0_PY, functions_from(PI0_PY, {"make_attn_mask", "posemb_sincos"}),
                            dict(jnp=jnp, jax=jax, einops=einops))
    exec_functions(PI0_PY, functions_from(PI0_PY, {"embed_suffix"}, cls="Pi0"), pi0_ns)
    pi0_ns["nnx"] = types.SimpleNamespace(swish=swish)
    met_ns = exec_functions(METRICS_PY, functions_from(METRICS_PY, {"compute_sample_specific_metrics"}), dict(jnp=jnp))
    fast_ns = exec_functions(PI0_FAST_PY, functions_from(PI0_FAST_PY, {"left_to_right_align", "put_along_last_axis"}),
                             dict(jnp=jnp, jax=jax))
    ns = dict(jnp=jnp, jax=jax, einops=einops, logger=logging.getLogger("openpi"), VQA_DATASET_ID_MAP={},
              _pi0=types.SimpleNamespace(make_attn_mask=pi0_ns["make_attn_mask"]),
              # (`@jax.vmap` on left_to_right_align is dropped with the other decorators and re-applied here)
              _pi0_fast=types.SimpleNamespace(left_to_right_align=jax.vmap(fast_ns["left_to_right_align"]),
                                              put_along_last_axis=fast_ns["put_along_last_axis"]),
              preprocess_observation=lambda rng, obs, **kw: obs,
              compute_sample_specific_metrics=met_ns["compute_sample_specific_metrics"],
              compute_per_vqa_dataset_metrics=None, compute_token_accuracy_metrics=None)
    exec_functions(LAP_PY, functions_from(LAP_PY, LAP_METHODS, cls="LAP"), ns)

    class RefLAP:
        pass

    for name in LAP_METHODS:
        setattr(RefLAP, name, ns[name])
    RefLAP.embed_suffix = pi0_ns["embed_suffix"]
    self = RefLAP()
    self._configure_shared_training_attributes(cfg)
    self.VOCAB_SIZE = cfg.vocab_size  # lap.py:33 hard-codes the PaliGemma vocabulary; the test model's is smaller
    self.EOS_TOKEN = 1                # lap.py:32
    self.action_horizon, self.action_dim = cfg.action_horizon, cfg.action_dim
    leaves = _Leaves(model, params["PaliGemma/llm/embedder/input_embedding"])
    self.PaliGemma = types.SimpleNamespace(img=leaves.img, llm=leaves.llm)
    for nm in ("action_in_proj", "action_out_proj", "time_mlp_in", "time_mlp_out"):
        setattr(self, nm, _linear(getattr(model, nm)))
    return self, leaves


class RefCoTObservation:
    """Attribute bag with the CoTObservation fields compute_loss / sample_actions read (model_adapter.py:37-80)."""

    def __init__(self, cfg, inp, langact=True):
        self.images = {k: inp["image/" + k] for k in cfg.image_keys}
        self.image_masks = {k: inp["image_mask/" + k] for k in cfg.image_keys}
        self.state = inp["state"]
        self.tokenized_prompt = inp["tokenized_prompt"]
        self.tokenized_prompt_mask = inp["tokenized_prompt_mask"]
        self.tokenized_langact_mask = inp["tokenized_langact_mask"] if langact else None
        self.token_loss_mask = inp["token_loss_mask"]
        self.sample_mask = inp["sample_mask"]
        self.is_vqa_sample = None
        self.is_prediction_sample = None
        self.critical_token_mask = self.number_token_mask = self.direction_token_mask = None


def run_case(pp, case):
    cfg = lap_case_config(case)
    batch, seed = LAP_CASES[case]["batch"], LAP_CASES[case]["seed"]
    params = seeded_reference_params(cfg, seed)
    inp = lap_case_inputs(cfg, batch, seed)
    ref, leaves = build_reference_lap(pp, cfg, params, inp["noise"], inp["time"])
    out = {"params_sha256": np.frombuffer(params_digest(params).encode(), dtype=np.uint8)}
    obs = RefCoTObservation(cfg, inp)
    pre_tok, pre_mask, pre_ar = ref.embed_prefix(obs)
    out["prefix_tokens"], out["prefix_mask"], out["prefix_ar_mask"] = pack_rows(pre_tok, G.ROW_STRIDE), pre_mask, pre_ar
    loss, metrics = ref.compute_loss(None, obs, inp["actions"], train=True)
    out["loss"] = np.float32(loss)
    for k in ("lang_loss", "langact_loss", "action_loss"):
        out[k] = np.float32(metrics[k])
    call = leaves.calls[-1]
    out["attn_mask"] = np.packbits(call["mask"], axis=-1)
    out["positions"] = call["positions"].astype(np.int32)
    # sample_actions as written: action rows see every valid prefix key, lang-action keys included
    for tag, langact in (("eval", True), ("serve", False)):
        o = RefCoTObservation(cfg, inp, langact=langact)
        out[f"sampled_actions_{tag}"] = np.asarray(ref.sample_actions(None, o, num_steps=10, noise=inp["noise"]),
                                                   dtype=np.float32)
    out["sampled_actions_serve_4"] = np.asarray(
        ref.sample_actions(None, RefCoTObservation(cfg, inp, langact=False), num_steps=4, noise=inp["noise"]), dtype=np.float32)
    # sample_tokens (lap.py:678-766), greedy: tokens + the logits every decode call produced (prefill logit first)
    S = 6
    leaves.decoded.clear()
    toks = ref.sample_tokens(None, RefCoTObservation(cfg, inp, langact=False), max_decoding_steps=S, temperature=0.0)
    out["ar_tokens"] = np.asarray(toks).astype(np.int32)
    out["ar_logits"] = np.concatenate([np.asarray(l, dtype=np.float32) for l in leaves.decoded], axis=1)  # [B, steps+1, V]
    # stand-alone helpers on awkward inputs (ragged validity, AR blocks)
    rng = np.random.default_rng(seed + 5)
    im = rng.random((3, 37)) < 0.8
    ar = rng.random((3, 37)) < 0.3
    out["kat_input_mask"], out["kat_ar_mask"] = im, ar
    out["kat_attn_mask"] = np.asarray(ref._build_combined_attention_mask(im, ar, im, None, None), dtype=bool)
    return out


def main():
    torch.manual_seed(0)
    pp = G.load_reference_pytorch_port()
    for case in LAP_CASES:
        out = run_case(pp, case)
        path = os.path.join(HERE, f"reference_lap_{case}.npz")
        np.savez_compressed(path, row_stride=np.int64(G.ROW_STRIDE), **out)
        print(case, {k: (v.shape if hasattr(v, "shape") and v.shape else v) for k, v in out.items() if k != "params_sha256"},
              os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
