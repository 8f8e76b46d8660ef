"""Generate tests/golden/reference_attention.npz: `Attention.__call__` of the reference's two-expert Gemma
(src/lap/models/backbones/gemma.py:167-290, with `_apply_rope` :548-564, `_init_cache` :586-594, `_name` :567-574), executed
from its source on seeded inputs with `stop_action_to_vlm_grad` False and True — forward outputs AND gradients.

The module is flax/JAX code; what the method body needs from those libraries is array algebra plus `stop_gradient`, so it is
run with TORCH standing in for jax.numpy (a Tensor subclass adds `.astype` and `.at[...].set`), `jax.lax.stop_gradient` ->
`detach()`, `jax.nn.softmax` -> `torch.softmax`, `lora.Einsum(shape, name, ...)` -> an einsum against the seeded weight of
that name (no LoRA in LAP-3B), `nn.compact` / typecheck decorators dropped.  Gradients then come from torch autograd through
the reference's own statements — in particular through its two `stop_gradient` sites (:248-253, :262-269), which is what pins
the N3 row (`lap` pre-training config).  fp32 throughout.  Run: python tests/golden/make_reference_attention_golden.py"""
import ast
import os
import sys
import types

import einops
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import reference_cases as RCm  # noqa: E402

GRAD_STRIDE = 8
REF = os.environ.get("LAP_REFERENCE", "/root/reference")
GEMMA_PY = os.path.join(REF, "src/lap/models/backbones/gemma.py")


class JT(torch.Tensor):
    """torch.Tensor with the two jax.Array members the executed code uses."""

    def astype(self, dtype):
        if isinstance(dtype, str):  # jax accepts dtype names ("bfloat16")
            dtype = getattr(torch, dtype)
        return self.to(dtype)

    @property
    def at(self):
        arr = self

        class _At:
            def __getitem__(self, idx):
                class _Set:
                    def set(self, value):
                        out = arr.clone()
                        out[idx] = value
                        return out
                return _Set()
        return _At()


def jt(x):
    return x.as_subclass(JT)


def _eq(eq):  # jax allows digits as einsum labels ("BSD,2KDH->2BSKH"); torch does not
    return eq.replace("2", "z").replace("3", "y")


class _Jnp:
    float32, int32 = torch.float32, torch.int32

    @staticmethod
    def concatenate(xs, axis=0):
        return torch.cat(list(xs), dim=axis)

    @staticmethod
    def full(shape, value, dtype=None):
        return jt(torch.full(tuple(shape), value, dtype=dtype))

    @staticmethod
    def zeros(shape, dtype=None):
        return jt(torch.zeros(tuple(shape), dtype=dtype))

    @staticmethod
    def arange(n, dtype=None):
        return jt(torch.arange(n, dtype=dtype))

    sin, cos = staticmethod(torch.sin), staticmethod(torch.cos)

    @staticmethod
    def split(x, n, axis=-1):
        return torch.chunk(x, n, dim=axis)

    bfloat16 = torch.bfloat16

    @staticmethod
    def einsum(eq, *ops, preferred_element_type=None):
        # XLA semantics for half-precision operands: exact products, fp32 accumulation, ONE rounding to the result type
        # (the operands' type unless preferred_element_type says otherwise)
        out = torch.einsum(_eq(eq), *[o.float() for o in ops])
        return out.to(preferred_element_type or ops[0].dtype)

    @staticmethod
    def where(c, a, b):
        b = b if isinstance(b, torch.Tensor) else torch.tensor(b, dtype=a.dtype)
        return torch.where(c, a, b)

    @staticmethod
    def pad(x, pad_width):
        flat = []
        for lo, hi in reversed(pad_width):
            flat += [lo, hi]
        return torch.nn.functional.pad(x, flat)


def load_reference_attention(params, layer, prefix="PaliGemma/llm/layers/attn/"):
    tree = ast.parse(open(GEMMA_PY).read())

    def strip(fn):
        fn.decorator_list, fn.returns = [], None
        for a in fn.args.args + fn.args.kwonlyargs:
            a.annotation = None
        return fn

    fns = [strip(n) for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("_apply_rope", "_init_cache", "_update_cache", "_name")]
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "Attention")
    call = strip(next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "__call__"))
    call.name = "attention_call"

    class Einsum:
        def __init__(self, shape, name, init_fn=None, lora_config=None):
            assert lora_config is None
            self.w = params[prefix + name + "/w"][layer]
            assert tuple(self.w.shape) == tuple(shape), (name, self.w.shape, shape)

        def __call__(self, eq, x):  # OP/models/lora.py:54-57: the fp32 parameter is cast to the activation dtype at use
            dtype = x.dtype
            return _Jnp.einsum(eq, x, self.w.astype(dtype))

    jax = types.SimpleNamespace(lax=types.SimpleNamespace(stop_gradient=lambda x: x.detach()),
                                nn=types.SimpleNamespace(softmax=lambda x, axis=-1: torch.softmax(x, dim=axis)))
    nn = types.SimpleNamespace(initializers=types.SimpleNamespace(lecun_normal=lambda **kw: None))
    ns = dict(jnp=_Jnp(), jax=jax, einops=einops, lora=types.SimpleNamespace(Einsum=Einsum), nn=nn)
    exec(compile(ast.fix_missing_locations(ast.Module(body=fns + [call], type_ignores=[])), GEMMA_PY, "exec"), ns)
    return ns["attention_call"]


def cases():
    from lap_b200.config import get_gemma_config
    out = {}
    for tag, (gn, en, B, P0, A, seed) in {"a": ("pin_a", "pin_a_expert", 2, 11, 5, 3), "b": ("pin_b", "pin_b_expert", 1, 7, 3, 4)}.items():
        g, e = get_gemma_config(gn), get_gemma_config(en)
        rng = np.random.default_rng(seed)
        N, H = g.num_heads, g.head_dim
        f = lambda *s: (rng.standard_normal(s) * 0.3).astype(np.float32)
        w = {"q_einsum/w": f(1, N, g.width, H), "kv_einsum/w": f(1, 2, 1, g.width, H), "attn_vec_einsum/w": f(1, N, H, g.width),
             "q_einsum_1/w": f(1, N, e.width, H), "kv_einsum_1/w": f(1, 2, 1, e.width, H), "attn_vec_einsum_1/w": f(1, N, H, e.width)}
        x0, x1 = f(B, P0, g.width), f(B, A, e.width)
        T = P0 + A
        valid = np.ones((B, T), bool)
        valid[0, 2] = False                                   # a masked prefix key
        ar = np.zeros((B, T), np.int32)
        ar[:, P0 - 3] = 1                                     # prefix-LM block boundary inside the prefix (lang-action tokens)
        ar[:, P0] = 1                                         # action tokens form their own block
        cs = np.cumsum(ar, 1)
        mask = (cs[:, None, :] <= cs[:, :, None]) & valid[:, None, :] & valid[:, :, None]   # pi0.make_attn_mask
        pos = np.cumsum(valid, 1) - 1
        out[tag] = dict(g=g, e=e, w=w, x0=x0, x1=x1, mask=mask, pos=pos.astype(np.int32), c0=f(B, P0, g.width), c1=f(B, A, e.width))
    return out


def main():
    res = {}
    for tag, c in cases().items():
        for stop in (False, True):
            leaf = lambda a: jt(torch.from_numpy(a.copy())).requires_grad_(True)
            params = {"PaliGemma/llm/layers/attn/" + k: leaf(v) for k, v in c["w"].items()}
            x0, x1 = leaf(c["x0"]), leaf(c["x1"])
            attention_call = load_reference_attention(params, 0)
            self = types.SimpleNamespace(configs=[types.SimpleNamespace(head_dim=k.head_dim, num_heads=k.num_heads, num_kv_heads=k.num_kv_heads,
                                                                        width=k.width, lora_configs={}) for k in (c["g"], c["e"])],
                                         stop_action_to_vlm_grad=stop, cache_dtype=None)
            mask = jt(torch.from_numpy(c["mask"]))[:, None]                      # [B, 1, T, S]
            out, (idx, k, v) = attention_call(self, [x0, x1], jt(torch.from_numpy(c["pos"])), mask, None)
            loss = (out[0] * torch.from_numpy(c["c0"])).sum() + (out[1] * torch.from_numpy(c["c1"])).sum()
            loss.backward()
            key = f"{tag}/stop{int(stop)}/"
            res[key + "out0"], res[key + "out1"] = out[0].detach().numpy(), out[1].detach().numpy()
            res[key + "gx0"], res[key + "gx1"] = x0.grad.numpy(), x1.grad.numpy()
            for n, p_ in params.items():  # weight gradients: every 8th element + (sum, norm, signed sum) of the whole tensor
                gw = p_.grad.numpy()
                res[key + "g/" + n.rsplit("attn/", 1)[1]] = gw.reshape(-1)[::GRAD_STRIDE].copy()
                res[key + "gf/" + n.rsplit("attn/", 1)[1]] = RCm.grad_fingerprint(gw)
        # bf16 activations (LAPConfig.dtype = "bfloat16"): the same source statements on bfloat16 inputs; forward values only.
        # torch follows the same promotion rules on these statements (bf16 x python scalar -> bf16, bf16 x fp32 -> fp32)
        for stop in (False, True):
            cast = lambda a: jt(torch.from_numpy(a.copy()).to(torch.bfloat16))
            params = {"PaliGemma/llm/layers/attn/" + k: jt(torch.from_numpy(v.copy())) for k, v in c["w"].items()}
            attention_call = load_reference_attention(params, 0)
            self = types.SimpleNamespace(configs=[types.SimpleNamespace(head_dim=k.head_dim, num_heads=k.num_heads, num_kv_heads=k.num_kv_heads,
                                                                        width=k.width, lora_configs={}) for k in (c["g"], c["e"])],
                                         stop_action_to_vlm_grad=stop, cache_dtype=None)
            with torch.no_grad():
                out, _ = attention_call(self, [cast(c["x0"]), cast(c["x1"])], jt(torch.from_numpy(c["pos"])),
                                        jt(torch.from_numpy(c["mask"]))[:, None], None)
            assert out[0].dtype == out[1].dtype == torch.bfloat16
            res[f"{tag}/bf16/stop{int(stop)}/out0"], res[f"{tag}/bf16/stop{int(stop)}/out1"] = out[0].float().numpy(), out[1].float().numpy()
        # (inputs and weights are regenerated from `cases()` by the test: pure numpy, no reference needed)
    np.savez_compressed(os.path.join(HERE, "reference_attention.npz"), **res)
    d = np.abs(res["a/stop1/gx0"] - res["a/stop0/gx0"]).max()
    print(len(res), "arrays; max |gx0(stop) - gx0(no stop)| =", d, "; out equal:", np.array_equal(res["a/stop1/out1"], res["a/stop0/out1"]))


if __name__ == "__main__":
    main()
