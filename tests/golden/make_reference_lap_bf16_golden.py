"""Generate tests/golden/reference_lap_bf16_*.npz: `LAP.compute_loss` and `LAP.sample_actions` in the precision the model
actually runs in (`LAPConfig.dtype = "bfloat16"`), from the reference's sources end to end:

  * `src/lap/models/lap.py` methods (compute_loss, embed_prefix, prepare_suffix, mask / loss helpers, sample_actions) and
    `pi0.py` (embed_suffix, make_attn_mask, posemb_sincos), executed from source on numpy arrays — bfloat16 arrays are real
    `ml_dtypes.bfloat16` numpy arrays, so the dtype flow of every statement (slicing, concatenation, promotion against fp32)
    is numpy's, which follows the same promotion lattice for these operations;
  * `self.PaliGemma.llm(...)`: the two-expert Gemma stack executed from `gemma.py` / `lora.py` source on torch bfloat16 tensors
    (make_reference_stack_golden.build: Module.__call__ -> Block -> RMSNorm / Attention / FeedForward), incl. the KV cache of
    sample_actions; `method="embed"` / `"decode"` follow `Embedder.encode` / `decode` + `Module.embed` (gemma.py:148-154,
    446-448): fp32 table row x fp32 sqrt(width), cast to bfloat16; bf16 activations x fp32 table -> fp32 logits;
  * the four `nnx.Linear` projections: fp32 parameters, inputs promoted to fp32 (flax promote_dtype);
  * `self.PaliGemma.img` (SigLIP): flax modules, not source-executable — the ORACLE's bf16 SigLIP is used for this leaf, so
    this fixture does not pin SigLIP (the fp32 SigLIP is pinned by the PyTorch port, reference_pi05_*.npz).
The oracle's `compute_loss(bf16=True)` / `sample_actions(bf16=True)` are compared with what this records.
Run: python tests/golden/make_reference_lap_bf16_golden.py"""
import math
import os
import sys
import types

import einops
import ml_dtypes
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import logging  # noqa: E402

import make_reference_lap_golden as L  # noqa: E402
import make_reference_stack_golden as S  # noqa: E402
import reference_cases as RC  # noqa: E402
from oracle import lap_oracle as O  # noqa: E402

BF = ml_dtypes.bfloat16


def to_np(t: torch.Tensor):
    return t.float().numpy().astype(BF) if t.dtype == torch.bfloat16 else t.numpy()


def to_torch(a):
    a = np.asarray(a)
    if a.dtype == BF:
        return S.jt(torch.from_numpy(a.astype(np.float32)).to(torch.bfloat16))
    return S.jt(torch.from_numpy(np.ascontiguousarray(a)))


def build(cfg, params, noise, time):
    tp = {k: torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)) for k, v in params.items()}
    run = S.build({k: v for k, v in tp.items() if k.startswith("PaliGemma/llm/")}, [cfg.gemma, cfg.expert], "bfloat16", False)
    table = params["PaliGemma/llm/embedder/input_embedding"].astype(np.float32)
    calls = []

    def img(image, train=False):
        with torch.no_grad():
            out = O.siglip_forward(tp, cfg, torch.from_numpy(np.ascontiguousarray(image, dtype=np.float32)), bf16=True)
        return out.numpy().astype(BF), None   # bf16 values held in fp32 -> exact

    def llm(embedded=None, *, method=None, positions=None, mask=None, adarms_cond=None, kv_cache=None):
        if method == "embed":   # Embedder.encode (fp32) then Module.embed's astype(embed_dtype)
            x = table[np.asarray(embedded)]
            x = x * np.float32(math.sqrt(table.shape[1]))
            return x.astype(BF)
        if method == "decode":  # jnp.dot(bf16, fp32) -> fp32
            return np.asarray(embedded).astype(np.float32) @ table.T
        assert method is None
        emb = [None if e is None else to_torch(e) for e in embedded]
        cond = [None if c is None else to_torch(np.asarray(c, dtype=np.float32)) for c in adarms_cond]
        with torch.no_grad():
            outs, cache = run(emb, to_torch(np.asarray(positions).astype(np.int32)), to_torch(np.asarray(mask, dtype=bool)), cond, kv_cache)
        calls.append(dict(mask=np.asarray(mask, dtype=bool)))
        return [None if o is None else to_np(o) for o in outs], cache

    # fp32 transcendental functions and matmuls go through torch, the library the oracle uses: a last-bit difference between
    # two libm / BLAS implementations in the fp32 time MLP would otherwise flip a bfloat16 rounding of the adaRMS condition
    # now and then and cascade through the stack (seen on case b: one flipped element, 7e-4 on the sampled actions)
    via_torch = lambda fn: (lambda x: fn(torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))).numpy())

    def linear(name):
        k, b = tp[name + "/kernel"], tp[name + "/bias"]

        def f(x):
            return (torch.from_numpy(np.ascontiguousarray(np.asarray(x).astype(np.float32))) @ k + b).numpy()
        f.out_features = k.shape[1]
        return f

    class PowViaTorch(np.ndarray):  # `python_float ** array` (posemb_sincos: periods down to 4e-3, i.e. sin arguments up to
        def __rpow__(self, base):   # ~1.6e3 - one ulp of the power moves the sine by 1e-4) evaluated by torch.pow like the oracle
            return torch.pow(torch.tensor(float(base), dtype=torch.float32), torch.from_numpy(np.asarray(self, dtype=np.float32))).numpy()

    class Jnp(L._Jnp):
        sin, cos, exp = staticmethod(via_torch(torch.sin)), staticmethod(via_torch(torch.cos)), staticmethod(via_torch(torch.exp))

        @staticmethod
        def linspace(a, b, n):
            return torch.linspace(a, b, n, dtype=torch.float32).numpy().view(PowViaTorch)

    jnp, jax = Jnp(), L.make_jax_shim(noise, time)
    swish = via_torch(lambda x: x * torch.sigmoid(x))
    pi0_ns = L.exec_functions(L.PI0_PY, L.functions_from(L.PI0_PY, {"make_attn_mask", "posemb_sincos"}), dict(jnp=jnp, jax=jax, einops=einops))
    L.exec_functions(L.PI0_PY, L.functions_from(L.PI0_PY, {"embed_suffix"}, cls="Pi0"), pi0_ns)
    pi0_ns["nnx"] = types.SimpleNamespace(swish=swish)
    met_ns = L.exec_functions(L.METRICS_PY, L.functions_from(L.METRICS_PY, {"compute_sample_specific_metrics"}), dict(jnp=jnp))
    methods = L.LAP_METHODS - {"sample_tokens"}
    ns = dict(jnp=jnp, jax=jax, einops=einops, logger=logging.getLogger("openpi"), VQA_DATASET_ID_MAP={},
              _pi0=types.SimpleNamespace(make_attn_mask=pi0_ns["make_attn_mask"]), preprocess_observation=lambda rng, obs, **kw: obs,
              compute_sample_specific_metrics=met_ns["compute_sample_specific_metrics"], compute_per_vqa_dataset_metrics=None,
              compute_token_accuracy_metrics=None)
    L.exec_functions(L.LAP_PY, L.functions_from(L.LAP_PY, methods, cls="LAP"), ns)

    class RefLAP:
        pass

    for name in methods:
        setattr(RefLAP, name, ns[name])
    RefLAP.embed_suffix = pi0_ns["embed_suffix"]
    self = RefLAP()
    self._configure_shared_training_attributes(cfg)
    self.VOCAB_SIZE, self.EOS_TOKEN = cfg.vocab_size, 1
    self.action_horizon, self.action_dim = cfg.action_horizon, cfg.action_dim
    self.PaliGemma = types.SimpleNamespace(img=img, llm=llm)
    for nm in ("action_in_proj", "action_out_proj", "time_mlp_in", "time_mlp_out"):
        setattr(self, nm, linear(nm))
    return self


def main():
    for case in RC.LAP_CASES:
        cfg = RC.lap_case_config(case)
        batch, seed = RC.LAP_CASES[case]["batch"], RC.LAP_CASES[case]["seed"]
        params = RC.seeded_reference_params(cfg, seed)
        inp = RC.lap_case_inputs(cfg, batch, seed)
        ref = build(cfg, params, inp["noise"], inp["time"])
        out = {"params_sha256": np.frombuffer(RC.params_digest(params).encode(), dtype=np.uint8)}
        obs = L.RefCoTObservation(cfg, inp)
        loss, metrics = ref.compute_loss(None, obs, inp["actions"], train=True)
        out["loss"] = np.float32(loss)
        for k in ("lang_loss", "langact_loss", "action_loss"):
            out[k] = np.float32(metrics[k])
        for tag, langact in (("eval", True), ("serve", False)):
            a = ref.sample_actions(None, L.RefCoTObservation(cfg, inp, langact=langact), num_steps=10, noise=inp["noise"])
            out[f"sampled_actions_{tag}"] = np.asarray(a, dtype=np.float32)
        path = os.path.join(HERE, f"reference_lap_bf16_{case}.npz")
        np.savez_compressed(path, **out)
        print(case, {k: (v.shape if getattr(v, "shape", ()) else v) for k, v in out.items() if k != "params_sha256"})


if __name__ == "__main__":
    main()
