"""Generate tests/golden/reference_stack.npz: the reference's two-expert Gemma STACK executed from its own source —
`Module.__call__` (src/lap/models/backbones/gemma.py:455-531) -> `Block.__call__` (:331-392) -> `RMSNorm.__call__`
(:113-131), `Attention.__call__` (:167-290), `_gated_residual` (:577-583), and `lora.Einsum.__call__` / `FeedForward.__call__`
/ `_dot` (third_party/openpi/src/openpi/models/lora.py:52-57,124-148) — on bfloat16 AND float32 activations, for the joint
prefix+suffix pass (training), a prefix-only pass and a suffix-only pass against the prefix KV cache (inference).

As in make_reference_attention_golden.py the method bodies are compiled from the source files and run with torch tensors
standing in for jax arrays.  What is NOT source (third-party flax / XLA behaviour, stated here and in the oracle header):
  * half-precision `einsum` / `dot`: exact products, fp32 accumulation, one rounding to the result dtype;
  * `nn.Dense(features, dtype=d)(x)`: x, kernel, bias cast to d; `dot` then `+ bias`, each producing a d array;
  * `nn.gelu` (tanh approximation) evaluated in fp32 on the half-precision input and rounded once (XLA keeps the
    intermediate of an elementwise fusion in fp32: xla_allow_excess_precision);
  * `nn.scan` over layers = a Python loop over the leading axis of the stacked parameters; `nn.remat`, sharding constraints,
    `sow` and dropout(0) are identities.
The oracle's bf16 mode must reproduce the bfloat16 results BIT FOR BIT (tests/test_reference_golden.py).
Run: python tests/golden/make_reference_stack_golden.py"""
import ast
import math
import os
import sys
import types

import einops
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import make_reference_attention_golden as A  # noqa: E402
import reference_cases as RC  # noqa: E402

LORA_PY = os.path.join(A.REF, "third_party/openpi/src/openpi/models/lora.py")
jt, _Jnp = A.jt, A._Jnp


class Jnp(_Jnp):
    square, sqrt, reciprocal = staticmethod(torch.square), staticmethod(torch.sqrt), staticmethod(torch.reciprocal)

    @staticmethod
    def mean(x, axis=-1, keepdims=False):
        return torch.mean(x, dim=axis, keepdim=keepdims)

    @staticmethod
    def asarray(x, dtype=None):
        return x if dtype is None else x.to(dtype)

    @staticmethod
    def dot(a, b):
        out = a.float() @ b.float()
        return out.to(a.dtype)

    @staticmethod
    def dtype(d):
        return {"bfloat16": torch.bfloat16, "float32": torch.float32}[d] if isinstance(d, str) else d


def _strip(fn):
    fn.decorator_list, fn.returns = [], None
    for a in fn.args.args + fn.args.kwonlyargs:
        a.annotation = None
    return fn


def _methods(path, cls, names):
    tree = ast.parse(open(path).read())
    body = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls).body
    got = {n.name: _strip(n) for n in body if isinstance(n, ast.FunctionDef) and n.name in names}
    assert set(got) == set(names), (cls, set(names) - set(got))
    return got


def _functions(path, names):
    tree = ast.parse(open(path).read())
    got = {n.name: _strip(n) for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names}
    assert set(got) == set(names)
    return got


def gelu_tanh(x):
    y = x.float()
    return (0.5 * y * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (y + 0.044715 * y ** 3)))).to(x.dtype)


def build(params, cfgs, embed_dtype, stop):
    """-> callable(embedded, positions, mask, adarms_cond, kv_cache) running Module.__call__ from source."""
    jnp = Jnp()
    state = {"layer": 0}
    LAY = "PaliGemma/llm/layers/"
    dt = lambda s: Jnp.dtype(s)

    def compile_into(ns, path, fns):
        exec(compile(ast.fix_missing_locations(ast.Module(body=list(fns), type_ignores=[])), path, "exec"), ns)

    # ---- lora.Einsum / lora.FeedForward (source) ----
    lora_ns = dict(jnp=jnp)
    ein = _methods(LORA_PY, "Einsum", ["__call__"])["__call__"]
    ein.name = "einsum_call"
    ff = _methods(LORA_PY, "FeedForward", ["__call__", "_dot"])
    ff["__call__"].name = "ff_call"
    compile_into(lora_ns, LORA_PY, [ein, ff["__call__"], ff["_dot"]])
    lora_ns["nn"] = types.SimpleNamespace(gelu=gelu_tanh)

    class Einsum:
        def __init__(self, shape, name, init_fn=None, lora_config=None):
            self.lora_config = lora_config
            self.w = jt(params[LAY + "attn/" + name + "/w"][state["layer"]])
            assert tuple(self.w.shape) == tuple(shape)

        def __call__(self, eqn, x):
            return lora_ns["einsum_call"](self, eqn, x)

    class FeedForward:
        def __init__(self, features, hidden_dim, name, lora_config=None):
            self.w_gating = jt(params[LAY + name + "/gating_einsum"][state["layer"]])
            self.w_linear = jt(params[LAY + name + "/linear"][state["layer"]])
            self.w_gating_lora = self.w_linear_lora = None
            assert tuple(self.w_gating.shape) == (2, features, hidden_dim)

        def _dot(self, x, w, lora_weights):
            return lora_ns["_dot"](self, x, w, lora_weights)

        def __call__(self, x):
            return lora_ns["ff_call"](self, x)

    # ---- gemma.py: RMSNorm, Attention, Block, Module (source) ----
    class Dense:  # flax.linen.Dense(features, dtype=d): promote x / kernel / bias to d, dot, add bias
        def __init__(self, features, kernel_init=None, dtype=None):
            self.dtype, self.features = dtype, features

        def __call__(self, x):
            k, b = Dense.scope_params
            assert k.shape[-1] == self.features
            x, k, b = x.to(self.dtype), jt(k).to(self.dtype), jt(b).to(self.dtype)
            return jnp.dot(x, k) + b

    nn = types.SimpleNamespace(initializers=types.SimpleNamespace(lecun_normal=lambda **kw: None, zeros_init=lambda: None, zeros=None),
                               Dense=Dense, Dropout=None)
    jax = types.SimpleNamespace(lax=types.SimpleNamespace(stop_gradient=lambda x: x.detach()),
                                nn=types.SimpleNamespace(softmax=lambda x, axis=-1: torch.softmax(x, dim=axis)),
                                tree=types.SimpleNamespace(map=lambda f, xs: [None if x is None else f(x) for x in xs]))
    sharding = types.SimpleNamespace(activation_sharding_constraint=lambda x: x)
    g_ns = dict(jnp=jnp, jax=jax, einops=einops, nn=nn, sharding=sharding, lora=types.SimpleNamespace(Einsum=Einsum, FeedForward=FeedForward))
    fns = _functions(A.GEMMA_PY, ["_apply_rope", "_init_cache", "_update_cache", "_name", "_gated_residual"])
    rms = _methods(A.GEMMA_PY, "RMSNorm", ["__call__"])["__call__"]
    rms.name = "rms_call"
    att = _methods(A.GEMMA_PY, "Attention", ["__call__"])["__call__"]
    att.name = "attention_call"
    blk = _methods(A.GEMMA_PY, "Block", ["__call__"])["__call__"]
    blk.name = "block_call"
    mod = _methods(A.GEMMA_PY, "Module", ["__call__"])["__call__"]
    mod.name = "module_call"
    compile_into(g_ns, A.GEMMA_PY, list(fns.values()) + [rms, att, blk, mod])

    class RMSNorm:
        def __init__(self, name):
            self.name = name
            self.scope = ("PaliGemma/llm/" if name.startswith("final_norm") else LAY) + name + "/"

        def param(self, pname, init, shape):
            w = params[self.scope + pname]
            return jt(w if self.name.startswith("final_norm") else w[state["layer"]])

        def __call__(self, x, cond):
            if cond is not None:
                k, b = params[self.scope + "Dense_0/kernel"], params[self.scope + "Dense_0/bias"]
                Dense.scope_params = (k, b) if self.name.startswith("final_norm") else (k[state["layer"]], b[state["layer"]])
            return g_ns["rms_call"](self, x, cond)

    class Attention:
        def __init__(self, configs, stop_action_to_vlm_grad, name, cache_dtype):
            self.configs, self.stop_action_to_vlm_grad, self.cache_dtype = configs, stop_action_to_vlm_grad, cache_dtype

        def __call__(self, xs, positions, attn_mask, kv_cache):
            return g_ns["attention_call"](self, xs, positions, attn_mask, kv_cache)

    g_ns.update(RMSNorm=RMSNorm, Attention=Attention)
    conf = [types.SimpleNamespace(head_dim=c.head_dim, num_heads=c.num_heads, num_kv_heads=c.num_kv_heads, width=c.width,
                                  mlp_dim=c.mlp_dim, depth=c.depth, lora_configs={}) for c in cfgs]
    block_self = types.SimpleNamespace(configs=conf, stop_action_to_vlm_grad=stop, cache_dtype=None, dropout=0.0, dropout_bdims=(),
                                       sow=lambda *a, **k: None)

    def layers(embedded, kv_cache, positions, mask, adarms_cond, deterministic):  # nn.scan(Block) over the stacked params
        caches = []
        for l in range(cfgs[0].depth):
            state["layer"] = l
            embedded, c = g_ns["block_call"](block_self, embedded, None if kv_cache is None else kv_cache[l], positions, mask,
                                             adarms_cond, deterministic)
            caches.append(c)
        return embedded, caches

    module_self = types.SimpleNamespace(configs=conf, embed_dtype=embed_dtype, layers=layers,
                                        final_norms=[RMSNorm(g_ns["_name"]("final_norm", i)) for i in range(2)])
    return lambda embedded, positions, mask, adarms_cond, kv_cache=None: g_ns["module_call"](
        module_self, embedded, positions, mask, adarms_cond, kv_cache=kv_cache)


def case_inputs(case):
    cfg = RC.lap_config(case)
    seed = RC.CASES[case][5]
    p = {k: v for k, v in RC.seeded_reference_params(cfg, seed).items() if k.startswith("PaliGemma/llm/")}
    rng = np.random.default_rng(seed + 77)
    B, P0, Asz = 2, 9, 4
    f = lambda *s: rng.standard_normal(s).astype(np.float32)
    x0, x1, cond = f(B, P0, cfg.gemma.width), f(B, Asz, cfg.expert.width), f(B, cfg.expert.width)
    T = P0 + Asz
    valid = np.ones((B, T), bool)
    valid[1, 3] = False
    ar = np.zeros((B, T), np.int32)
    ar[:, P0 - 2] = 1
    ar[:, P0] = 1
    cs = np.cumsum(ar, 1)
    mask = (cs[:, None, :] <= cs[:, :, None]) & valid[:, None, :] & valid[:, :, None]
    pos = (np.cumsum(valid, 1) - 1).astype(np.int32)
    return cfg, p, dict(x0=x0, x1=x1, cond=cond, mask=mask, pos=pos, P0=P0)


def main():
    res = {}
    for case in RC.CASES:
        cfg, p, inp = case_inputs(case)
        tp = {k: torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)) for k, v in p.items()}
        P0 = inp["P0"]
        for dname in ("bfloat16", "float32"):
            for stop in (False, True):
                run = build(tp, [cfg.gemma, cfg.expert], dname, stop)
                t = lambda a: jt(torch.from_numpy(a.copy()))
                with torch.no_grad():
                    # joint pass (training): embedded arrive in fp32 and are cast to embed_dtype by Module.__call__
                    (o0, o1), _ = run([t(inp["x0"]), t(inp["x1"])], t(inp["pos"]), t(inp["mask"]), [None, t(inp["cond"])])
                    key = f"{case}/{dname}/stop{int(stop)}/"
                    res[key + "joint0"], res[key + "joint1"] = o0.float().numpy(), o1.float().numpy()
                    if not stop:
                        # prefix-only pass, then the suffix against its cache (inference, lap.py:634-665)
                        (q0, _), cache = run([t(inp["x0"]), None], t(inp["pos"][:, :P0]), t(inp["mask"][:, :P0, :P0]), [None, None])
                        res[key + "prefix0"] = q0.float().numpy()
                        kv = [(c[0], c[1], c[2]) for c in cache]
                        (_, s1), _ = run([None, t(inp["x1"])], t(inp["pos"][:, P0:]), t(inp["mask"][:, P0:, :]), [None, t(inp["cond"])], kv_cache=kv)
                        res[key + "suffix1"] = s1.float().numpy()
        for k in ("x0", "x1", "cond", "mask", "pos"):
            res[f"{case}/{k}"] = inp[k]
        res[f"{case}/P0"] = np.int64(P0)
    np.savez_compressed(os.path.join(HERE, "reference_stack.npz"), **res)
    print(len(res), "arrays", os.path.getsize(os.path.join(HERE, "reference_stack.npz")), "bytes")


if __name__ == "__main__":
    main()
