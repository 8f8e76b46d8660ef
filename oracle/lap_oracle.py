"""ORACLE — CPU restatement (PyTorch eager, fp32) of the reference's LAP-3B hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import it; nothing under lap_b200/ does.

PARITY PIN (what the restatement has been checked against, tests/test_reference_golden.py):
  * the reference's own PyTorch port of the π0.5 arithmetic LAP shares (OP/models_pytorch/pi0_pytorch.py,
    gemma_pytorch.py, transformers_replace/**), run UNMODIFIED in the build container on seeded weights mapped by the
    reference's JAX->PyTorch converter: SigLIP tower, token embedding, suffix embedding / adaRMS condition,
    make_attn_mask + positions (bit-exact), the joint two-expert Gemma stack, the flow-matching squared error and
    sample_actions (KV cache + Euler loop) agree with the fp32 mode of this file to <= 3e-5 normwise, and the
    gradients of the flow-matching loss from the port's autograd (scattered back to the JAX layout by applying the
    converter to an index tree) agree with autograd through this file to <= 1e-3 per tensor
    (fixtures tests/golden/reference_pi05_*.npz, generator tests/golden/make_reference_golden.py);
  * LAP.compute_loss / embed_prefix / prepare_suffix / the lang-action mask builders / the language CE and loss
    weighting / sample_actions / sample_tokens (greedy), executed from src/lap/models/lap.py's own source with numpy
    standing in for jax.numpy and the PyTorch port as leaf modules: masks and positions bit-exact, loss and metrics to
    <= 2e-4, decoded tokens identical and per-step logits to <= 2e-4 incl. the all-EOS early stop
    (fixtures tests/golden/reference_lap_*.npz, generator tests/golden/make_reference_lap_golden.py);
  * the two-expert Gemma stack executed from the reference's own source under torch (tensors standing in for jax arrays):
    gemma.Module.__call__ -> Block -> RMSNorm / Attention / _gated_residual (src/lap/models/backbones/gemma.py) and
    lora.Einsum / FeedForward (OP/models/lora.py), on bfloat16 and float32 activations, joint / prefix-only / suffix-with-
    cache passes, with and without stop_action_to_vlm_grad: the bf16 mode of this file reproduces the bfloat16 results BIT
    FOR BIT (a single dropped rounding changes 74 % of the outputs), fp32 agrees to 3e-7, and the attention block's
    gradients — including the two stop_gradient sites — agree with autograd through the reference statements to 2e-5
    (fixtures tests/golden/reference_stack.npz, reference_attention.npz; generators make_reference_stack_golden.py,
    make_reference_attention_golden.py, which also list the flax/XLA behaviours assumed: half-precision einsum = fp32
    accumulation + one rounding, nn.Dense(dtype=bf16) = dot then bias add, gelu evaluated in fp32 and rounded once);
  * end to end in bfloat16: LAP.compute_loss and sample_actions from lap.py / pi0.py source on real bfloat16 (ml_dtypes)
    arrays with the Gemma stack from source underneath (SigLIP leaf = this file's own): losses agree to fp32 round-off and
    the sampled actions (prefill + ten Euler steps through the KV cache) are BIT-IDENTICAL to the bf16 mode of this file
    (fixtures reference_lap_bf16_*.npz, generator make_reference_lap_bf16_golden.py).  A finding of that exercise:
    posemb_sincos (pi0.py:47-63) is ill-conditioned in fp32 — periods down to 4e-3 make one ulp of `pow` move the sine by
    1e-4 — so two correct implementations of the time embedding differ by ~1e-5 in the adaRMS condition, which flips an
    occasional bfloat16 rounding and moves sampled actions by ~7e-4 normwise: that is the floor any bf16 parity claim has;
  * the three worked `make_attn_mask` examples of OP/models/pi0.py:26-33 and structural invariants
    (tests/test_oracle.py).
STILL UNPINNED (the JAX/Flax program itself cannot run here: no jax/flax/optax wheels, no network; the reference's
tests hold no numeric vector): where XLA's fusions keep excess precision relative to the source's dtype flow (the dtype
flow of the Gemma stack and of lap.py is pinned above; SigLIP's flax modules in bf16 follow SURVEY.md Appendix A by
reading the source, their fp32 arithmetic is pinned by the PyTorch port), and the optax/EMA train-step arithmetic (third-party optax, restated from its
published definitions; checked against closed forms and against torch.optim.AdamW as an independent implementation).  Every function cites the file:line it follows
(`OP/` = third_party/openpi/src/openpi/).

Two precision modes:
  bf16=False : everything in fp32 (the "mathematical" function).
  bf16=True  : a rounding to bfloat16 is inserted exactly where the JAX code produces a bf16 array
               (SURVEY.md Appendix A): weights cast at use, activations rounded after every op that the
               reference evaluates in bf16, fp32 kept where the reference keeps it (norm statistics, attention
               logits/softmax, losses, suffix embedding).
All tensors are torch.float32 on CPU; `r()` rounds through bfloat16.  Gradients come from torch autograd
(casts are straight-through, as in JAX).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

BIG_NEG = -2.3819763e38  # gemma.py:258


def _r(x: torch.Tensor, bf16: bool) -> torch.Tensor:
    """Round to bfloat16 and back (identity in fp32 mode). Straight-through for autograd, like a JAX cast."""
    if not bf16:
        return x
    return x.to(torch.bfloat16).to(torch.float32)


# ---------------------------------------------------------------------------------------------
# masks / positions  (integer & bool work: must be bit-exact)
# ---------------------------------------------------------------------------------------------
def make_attn_mask(input_mask: torch.Tensor, mask_ar: torch.Tensor) -> torch.Tensor:
    """OP/models/pi0.py:19-44."""
    mask_ar = mask_ar.expand(input_mask.shape)
    cumsum = torch.cumsum(mask_ar.to(torch.int64), dim=1)
    attn = cumsum[:, None, :] <= cumsum[:, :, None]
    valid = input_mask[:, None, :] & input_mask[:, :, None]
    return attn & valid


def build_prefix_action_mask(prefix_mask, langact_mask):
    """lap.py:303-325."""
    if langact_mask is None:
        return prefix_mask
    img_len = prefix_mask.shape[1] - langact_mask.shape[1]
    full = torch.cat([torch.zeros(langact_mask.shape[0], img_len, dtype=torch.bool), langact_mask], dim=1)
    return prefix_mask & ~full


def build_combined_attention_mask(prefix_mask, prefix_ar_mask, prefix_mask_action, suffix_mask, suffix_ar_mask):
    """lap.py:327-364."""
    prefix_attn = make_attn_mask(prefix_mask, prefix_ar_mask)
    if suffix_mask is None:
        return prefix_attn
    b, p = prefix_mask.shape
    s = suffix_mask.shape[1]
    combined = torch.zeros(b, p + s, p + s, dtype=torch.bool)
    combined[:, :p, :p] = prefix_attn
    input_mask = torch.cat([prefix_mask_action, suffix_mask], dim=1)
    ar_mask = torch.cat([torch.zeros_like(prefix_mask_action), suffix_ar_mask], dim=1)
    action_mask = make_attn_mask(input_mask, ar_mask)
    combined[:, p:, :] = action_mask[:, p:, :]
    return combined


def build_combined_positions(prefix_mask, prefix_mask_action, suffix_mask):
    """lap.py:366-377."""
    pp = torch.cumsum(prefix_mask.to(torch.int64), dim=1) - 1
    if suffix_mask is None:
        return pp.to(torch.int32)
    sp = prefix_mask_action.sum(-1, keepdim=True) + torch.cumsum(suffix_mask.to(torch.int64), dim=-1) - 1
    return torch.cat([pp, sp], dim=1).to(torch.int32)


# ---------------------------------------------------------------------------------------------
# SigLIP  (OP/models/siglip.py; flax 0.10.2 semantics for LayerNorm / Dense / MHA restated)
# ---------------------------------------------------------------------------------------------
def layer_norm(x, scale, bias, bf16):
    """flax nn.LayerNorm(dtype=bf16): statistics in fp32 with the fast-variance formula, eps=1e-6, output bf16.
    siglip.py:87,98,161."""
    mean = x.mean(-1, keepdim=True)
    mean2 = (x * x).mean(-1, keepdim=True)
    var = torch.clamp(mean2 - mean * mean, min=0.0)
    y = (x - mean) * (torch.rsqrt(var + 1e-6) * scale) + bias
    return _r(y, bf16)


def dense(x, kernel, bias, bf16):
    """flax nn.Dense(dtype=bf16): y = bf16(bf16(x)@bf16(W)); y = bf16(y + bf16(b)).  siglip.py:69-72,286."""
    y = _r(x @ _r(kernel, bf16), bf16)
    if bias is not None:
        y = _r(y + _r(bias, bf16), bf16)
    return y


def gelu_tanh(x):
    return F.gelu(x, approximate="tanh")


def siglip_forward(p: dict, cfg, image: torch.Tensor, *, bf16: bool, softmax_dtype: str = "bf16",
                   prefix: str = "PaliGemma/img/") -> torch.Tensor:
    """siglip._Module.__call__ (siglip.py:208-290) with pool_type='none', scan=True.  image [N,H,W,3] fp32 in [-1,1]."""
    sc = cfg.siglip
    N, H, W, _ = image.shape
    ps = sc.patch_size
    gh, gw = H // ps, W // ps
    # fp32 patch conv, VALID, stride=patch (siglip.py:216-223); kernel [ps,ps,3,width]
    patches = image[:, : gh * ps, : gw * ps].reshape(N, gh, ps, gw, ps, 3).permute(0, 1, 3, 2, 4, 5).reshape(N, gh * gw, ps * ps * 3)
    wk = p[prefix + "embedding/kernel"].reshape(ps * ps * 3, sc.width)
    x = patches @ wk + p[prefix + "embedding/bias"]
    x = x + p[prefix + "pos_embedding"]  # fp32 (siglip.py:229)
    x = _r(x, bf16)  # cast to dtype_mm (siglip.py:239)
    blk = prefix + "Transformer/encoderblock/"
    hd = sc.head_dim
    nh = sc.num_heads
    for l in range(sc.depth):
        y = layer_norm(x, p[blk + "LayerNorm_0/scale"][l], p[blk + "LayerNorm_0/bias"][l], bf16)
        mha = blk + "MultiHeadDotProductAttention_0/"
        q = dense(y, p[mha + "query/kernel"][l].reshape(sc.width, nh * hd), p[mha + "query/bias"][l].reshape(-1), bf16)
        k = dense(y, p[mha + "key/kernel"][l].reshape(sc.width, nh * hd), p[mha + "key/bias"][l].reshape(-1), bf16)
        v = dense(y, p[mha + "value/kernel"][l].reshape(sc.width, nh * hd), p[mha + "value/bias"][l].reshape(-1), bf16)
        q = q.reshape(N, -1, nh, hd)
        k = k.reshape(N, -1, nh, hd)
        v = v.reshape(N, -1, nh, hd)
        # flax dot_product_attention: query / sqrt(depth).astype(dtype)
        depth_scale = _r(torch.tensor(math.sqrt(hd), dtype=torch.float32), bf16)
        q = _r(q / depth_scale, bf16)
        logits = _r(torch.einsum("nqhd,nkhd->nhqk", q, k), bf16)
        if bf16 and softmax_dtype == "bf16":
            # jax.nn.softmax evaluated in bf16 (flax force_fp32_for_softmax=False)
            m = logits.max(-1, keepdim=True).values
            e = _r(torch.exp(_r(logits - m, True)), True)
            w = _r(e / _r(e.sum(-1, keepdim=True), True), True)
        else:
            w = _r(torch.softmax(logits, dim=-1), bf16)
        o = _r(torch.einsum("nhqk,nkhd->nqhd", w, v), bf16).reshape(N, -1, nh * hd)
        y = dense(o, p[mha + "out/kernel"][l].reshape(nh * hd, sc.width), p[mha + "out/bias"][l], bf16)
        x = _r(x + y, bf16)
        y = layer_norm(x, p[blk + "LayerNorm_1/scale"][l], p[blk + "LayerNorm_1/bias"][l], bf16)
        h = dense(y, p[blk + "MlpBlock_0/Dense_0/kernel"][l], p[blk + "MlpBlock_0/Dense_0/bias"][l], bf16)
        h = _r(gelu_tanh(h), bf16)
        y = dense(h, p[blk + "MlpBlock_0/Dense_1/kernel"][l], p[blk + "MlpBlock_0/Dense_1/bias"][l], bf16)
        x = _r(x + y, bf16)
    x = layer_norm(x, p[prefix + "Transformer/encoder_norm/scale"], p[prefix + "Transformer/encoder_norm/bias"], bf16)
    return dense(x, p[prefix + "head/kernel"], p[prefix + "head/bias"], bf16)


# ---------------------------------------------------------------------------------------------
# Gemma multi-expert transformer  (src/lap/models/backbones/gemma.py)
# ---------------------------------------------------------------------------------------------
def _name(name, i):
    return name if i == 0 else f"{name}_{i}"  # gemma.py:567-574


def rms_norm(x, p, key, layer, cond, bf16):
    """gemma.py:112-131.  Returns (normed, gate)."""
    var = (x * x).mean(-1, keepdim=True)
    normed = x * torch.rsqrt(var + 1e-6)  # fp32
    if cond is None:
        scale = p[key + "/scale"] if layer is None else p[key + "/scale"][layer]
        return _r(normed * (1.0 + scale), bf16), None
    kern = p[key + "/Dense_0/kernel"] if layer is None else p[key + "/Dense_0/kernel"][layer]
    bias = p[key + "/Dense_0/bias"] if layer is None else p[key + "/Dense_0/bias"][layer]
    mod = dense(_r(cond, bf16), kern, bias, bf16)  # nn.Dense(dtype=bf16)
    scale, shift, gate = torch.chunk(mod[:, None, :], 3, dim=-1)
    one_plus = _r(1.0 + scale, bf16)  # bf16 arithmetic on a bf16 array
    return _r(normed * one_plus + shift, bf16), gate


def apply_rope(x, positions, bf16, max_wavelength=10_000):
    """gemma.py:548-564. x [B,L,H,D], positions [B,L] int."""
    d = x.shape[-1]
    freq_exponents = (2.0 / d) * torch.arange(d // 2, dtype=torch.float32)
    timescale = max_wavelength ** freq_exponents
    radians = positions[..., None].to(torch.float32) / timescale[None, None, :]
    radians = radians[..., None, :]
    sin, cos = torch.sin(radians), torch.cos(radians)
    x1, x2 = torch.chunk(x, 2, dim=-1)
    res = torch.cat([x1 * cos - x2 * sin, x2 * cos + x1 * sin], dim=-1)
    return _r(res, bf16)


def gated_residual(x, y, gate, bf16):
    """gemma.py:577-583 (bf16 arithmetic)."""
    if gate is None:
        return _r(x + y, bf16)
    return _r(x + _r(y * gate, bf16), bf16)


def gemma_attention(p, cfgs, layer, xs, positions, attn_mask, kv_cache, bf16, pre="PaliGemma/llm/layers/attn/",
                    stop_action_to_vlm_grad=False):
    """gemma.py:167-290.  xs: list per expert of [B,T_i,D_i] or None; attn_mask [B,T,S] bool; kv_cache (k,v) or None.
    stop_action_to_vlm_grad (gemma.py:206-213,242-269): queries of experts > 0 see expert 0's K and V through
    stop_gradient — the forward value is the same function, evaluated as two value products that are rounded separately."""
    qs, ks, vs = [], [], []
    for i, (x, c) in enumerate(zip(xs, cfgs)):
        if x is None:
            continue
        wq = _r(p[pre + _name("q_einsum", i) + "/w"][layer], bf16)  # [N,D,H]
        wkv = _r(p[pre + _name("kv_einsum", i) + "/w"][layer], bf16)  # [2,K,D,H]
        qs.append(_r(torch.einsum("btd,ndh->btnh", x, wq), bf16))
        kv = _r(torch.einsum("bsd,ckdh->cbskh", x, wkv), bf16)
        ks.append(kv[0])
        vs.append(kv[1])
    q, k, v = torch.cat(qs, 1), torch.cat(ks, 1), torch.cat(vs, 1)
    hd = cfgs[0].head_dim
    q = apply_rope(q, positions, bf16)
    q = _r(q * (hd ** -0.5), bf16)
    k = apply_rope(k, positions, bf16)
    if kv_cache is not None:
        ck, cv = kv_cache
        k = torch.cat([ck, k], 1)  # gemma.py:227-230 (suffix-only pass)
        v = torch.cat([cv, v], 1)
    B, T, Nh, _ = q.shape
    K = cfgs[0].num_kv_heads
    qg = q.reshape(B, T, K, Nh // K, hd)
    logits = torch.einsum("btkgh,bskh->bkgts", qg, k)  # fp32
    assert attn_mask.shape == (B, T, k.shape[1]), (attn_mask.shape, q.shape, k.shape)
    stop = stop_action_to_vlm_grad and xs[0] is not None and any(x is not None for x in xs[1:])
    if stop:
        P0 = xs[0].shape[1]  # expert-0 tokens come first (gemma.py:204)
        logits0_i = torch.einsum("btkgh,bskh->bkgts", qg[:, P0:], k[:, :P0].detach())  # gemma.py:248-253
        logits = torch.cat([logits[:, :, :, :P0], torch.cat([logits0_i, logits[:, :, :, P0:, P0:]], -1)], 3)
    masked = torch.where(attn_mask[:, None, None, :, :], logits, torch.tensor(BIG_NEG))
    probs = _r(torch.softmax(masked, dim=-1), bf16)
    if stop:
        cross = torch.zeros(T, k.shape[1], dtype=torch.bool)
        cross[P0:, :P0] = True
        probs_cross = probs * cross.to(probs.dtype)
        probs_self = probs - probs_cross
        enc = _r(_r(torch.einsum("bkgts,bskh->btkgh", probs_self, v), bf16)
                 + _r(torch.einsum("bkgts,bskh->btkgh", probs_cross, v.detach()), bf16), bf16).reshape(B, T, Nh, hd)
    else:
        enc = _r(torch.einsum("bkgts,bskh->btkgh", probs, v), bf16).reshape(B, T, Nh, hd)
    out, start = [], 0
    for i, (x, c) in enumerate(zip(xs, cfgs)):
        if x is None:
            out.append(None)
            continue
        end = start + x.shape[1]
        wo = _r(p[pre + _name("attn_vec_einsum", i) + "/w"][layer], bf16)  # [N,H,D]
        out.append(_r(torch.einsum("btnh,nhd->btd", enc[:, start:end], wo), bf16))
        start = end
    return out, (k, v)


def feed_forward(p, key, layer, x, bf16):
    """OP/models/lora.py:124-148 (GeGLU)."""
    wg = _r(p[key + "/gating_einsum"][layer], bf16)  # [2,D,F]
    g = _r(x @ wg[0], bf16)
    u = _r(x @ wg[1], bf16)
    act = _r(_r(gelu_tanh(g), bf16) * u, bf16)
    return _r(act @ _r(p[key + "/linear"][layer], bf16), bf16)


def gemma_forward(p, cfgs, embedded, positions, mask, adarms_cond, bf16, kv_cache=None,
                  pre="PaliGemma/llm/", stop_action_to_vlm_grad=False):
    """gemma.Module.__call__ (gemma.py:455-531). Returns (outputs per expert, per-layer kv list)."""
    xs = [None if e is None else _r(e, bf16) for e in embedded]  # astype(embed_dtype), gemma.py:494
    depth = cfgs[0].depth
    new_cache = []
    lay = pre + "layers/"
    for l in range(depth):
        pre_attn, gates = [], []
        for i, x in enumerate(xs):
            if x is None:
                pre_attn.append(None)
                gates.append(None)
                continue
            h, g = rms_norm(x, p, lay + _name("pre_attention_norm", i), l, adarms_cond[i], bf16)
            pre_attn.append(h)
            gates.append(g)
        post, kv = gemma_attention(p, cfgs, l, pre_attn, positions, mask, None if kv_cache is None else kv_cache[l],
                                   bf16, pre=lay + "attn/", stop_action_to_vlm_grad=stop_action_to_vlm_grad)
        new_cache.append(kv)
        xs = [None if x is None else gated_residual(x, y, g, bf16) for x, y, g in zip(xs, post, gates)]
        outs, gates = [], []
        for i, x in enumerate(xs):
            if x is None:
                outs.append(None)
                gates.append(None)
                continue
            h, g = rms_norm(x, p, lay + _name("pre_ffw_norm", i), l, adarms_cond[i], bf16)
            outs.append(feed_forward(p, lay + _name("mlp", i), l, h, bf16))
            gates.append(g)
        xs = [None if x is None else gated_residual(x, y, g, bf16) for x, y, g in zip(xs, outs, gates)]
    final = []
    for i, x in enumerate(xs):
        if x is None:
            final.append(None)
            continue
        final.append(rms_norm(x, p, pre + _name("final_norm", i), None, adarms_cond[i], bf16)[0])
    return final, new_cache


def embed_tokens(p, cfg, tokens, bf16):
    """Embedder.encode + Module.embed (gemma.py:148-151,446-448)."""
    E = p["PaliGemma/llm/embedder/input_embedding"]
    x = E[tokens.long()] * torch.sqrt(torch.tensor(float(cfg.gemma.width)))
    return _r(x, bf16)


# ---------------------------------------------------------------------------------------------
# suffix (flow matching) — all fp32 until gemma.py:494
# ---------------------------------------------------------------------------------------------
def posemb_sincos(pos, dim, min_period, max_period):
    """OP/models/pi0.py:47-63."""
    fraction = torch.linspace(0.0, 1.0, dim // 2, dtype=torch.float32)
    period = min_period * (max_period / min_period) ** fraction
    inp = pos[:, None].to(torch.float32) * (1.0 / period * 2 * math.pi)[None, :]
    return torch.cat([torch.sin(inp), torch.cos(inp)], dim=-1)


def swish(x):
    return x * torch.sigmoid(x)


def embed_suffix(p, cfg, noisy_actions, timestep):
    """Pi0.embed_suffix, pi05 branch (OP/models/pi0.py:139-186). Returns tokens, mask, ar_mask[S], adarms_cond."""
    assert cfg.pi05
    tok = noisy_actions @ p["action_in_proj/kernel"] + p["action_in_proj/bias"]
    width = p["action_in_proj/kernel"].shape[1]
    te = posemb_sincos(timestep, width, 4e-3, 4.0)
    te = swish(te @ p["time_mlp_in/kernel"] + p["time_mlp_in/bias"])
    te = swish(te @ p["time_mlp_out/kernel"] + p["time_mlp_out/bias"])
    B, A = noisy_actions.shape[:2]
    mask = torch.ones(B, A, dtype=torch.bool)
    ar = torch.tensor([True] + [False] * (A - 1))
    return tok, mask, ar, te


# ---------------------------------------------------------------------------------------------
# LAP model  (src/lap/models/lap.py)
# ---------------------------------------------------------------------------------------------
def embed_prefix(p, cfg, obs, bf16, softmax_dtype="bf16"):
    """lap.py:118-170. obs: dict with images{name:[B,H,W,3]}, image_masks{name:[B]}, tokenized_prompt,
    tokenized_prompt_mask, tokenized_langact_mask (or None)."""
    toks, masks, ars = [], [], []
    for name in cfg.image_keys:
        it = siglip_forward(p, cfg, obs["images"][name], bf16=bf16, softmax_dtype=softmax_dtype)
        toks.append(it)
        masks.append(obs["image_masks"][name][:, None].expand(-1, it.shape[1]))
        ars.append(torch.zeros(it.shape[0], it.shape[1], dtype=torch.bool))
    toks.append(embed_tokens(p, cfg, obs["tokenized_prompt"], bf16))
    masks.append(obs["tokenized_prompt_mask"])
    la = obs.get("tokenized_langact_mask")
    ars.append(la if la is not None else torch.zeros_like(obs["tokenized_prompt_mask"]))
    return torch.cat(toks, 1), torch.cat(masks, 1), torch.cat(ars, 1)


def decode_logits(p, x, bf16):
    """Embedder.decode (gemma.py:153-154): bf16 activations times the fp32 table -> fp32 logits."""
    return x @ p["PaliGemma/llm/embedder/input_embedding"].T


def compute_loss(p, cfg, obs, actions, noise, time, *, bf16: bool, softmax_dtype="bf16", return_aux=False):
    """LAP.compute_loss (lap.py:380-602) for enable_langact_training & enable_action_training, no VQA/prediction
    (the lap_libero configuration).  `noise`/`time` replace jax.random (lap.py:193-194)."""
    B = actions.shape[0]
    # prepare_suffix (lap.py:185-207)
    te = time[:, None, None]
    x_t = te * noise + (1 - te) * actions
    u_t = noise - actions
    suf_tok, suf_mask, suf_ar, cond = embed_suffix(p, cfg, x_t, time)
    suf_ar = suf_ar[None, :].expand(B, -1)
    pre_tok, pre_mask, pre_ar = embed_prefix(p, cfg, obs, bf16, softmax_dtype)
    la = obs.get("tokenized_langact_mask")
    pre_mask_action = build_prefix_action_mask(pre_mask, la)
    mask = build_combined_attention_mask(pre_mask, pre_ar, pre_mask_action, suf_mask, suf_ar)
    positions = build_combined_positions(pre_mask, pre_mask_action, suf_mask)
    (pre_out, suf_out), _ = gemma_forward(p, [cfg.gemma, cfg.expert], [pre_tok, suf_tok], positions, mask,
                                          [None, cond], bf16,
                                          stop_action_to_vlm_grad=getattr(cfg, "stop_action_to_vlm_grad", False))
    metrics = {}
    # language loss (lap.py:209-289)
    tp = obs["tokenized_prompt"].long()
    L = tp.shape[1]
    targets = tp[:, 1:]
    pl = pre_out[:, :-1][:, -(L - 1):]
    logits = decode_logits(p, pl, bf16)
    loss_mask = la[:, 1:] & obs["tokenized_prompt_mask"][:, 1:] & obs["token_loss_mask"][:, 1:]
    sample_mask = obs.get("sample_mask")
    lm = loss_mask.to(torch.float32)
    if sample_mask is not None:
        lm = lm * sample_mask[:, None].to(torch.float32)
    logp = torch.log_softmax(logits, dim=-1)
    tok_ll = torch.gather(logp, 2, targets[..., None])[..., 0]
    lang_per_sample = -(tok_ll * lm).sum(-1) / torch.clamp(lm.sum(-1), min=1.0)
    metrics["lang_loss"] = lang_per_sample.mean()
    # compute_sample_specific_metrics(prefix="langact_") -> masked mean (model_utils/metrics.py)
    if sample_mask is not None:
        smf = sample_mask.to(torch.float32)
        metrics["langact_loss"] = (lang_per_sample * smf).sum() / torch.clamp(smf.sum(), min=1.0)
    lang_w = cfg.language_loss_weight * lang_per_sample
    # action loss (lap.py:291-301)
    A = cfg.action_horizon
    v_t = suf_out[:, -A:] @ p["action_out_proj/kernel"] + p["action_out_proj/bias"]
    action_per_sample = ((v_t - u_t) ** 2).mean(dim=(-1, -2))
    metrics["action_loss"] = action_per_sample.mean()
    act_w = cfg.action_loss_weight * action_per_sample
    # final (lap.py:573-596)
    action_term = act_w.sum() / max(float(B), 1.0)
    if sample_mask is not None:
        lang_term = lang_w.sum() / torch.clamp(sample_mask.to(torch.float32).sum(), min=1.0)
    else:
        lang_term = lang_w.mean()
    loss = lang_term + action_term
    if return_aux:
        return loss, metrics, dict(prefix_out=pre_out, suffix_out=suf_out, mask=mask, positions=positions,
                                   prefix_tokens=pre_tok, v_t=v_t, logits=logits, lang_per_sample=lang_per_sample,
                                   action_per_sample=action_per_sample, cond=cond, suffix_tokens=suf_tok)
    return loss, metrics


def sample_actions(p, cfg, obs, noise, *, num_steps=10, bf16: bool, softmax_dtype="bf16"):
    """LAP.sample_actions (lap.py:605-675): prefix pass -> KV cache -> num_steps Euler steps."""
    dt = -1.0 / num_steps
    B = noise.shape[0]
    pre_tok, pre_mask, pre_ar = embed_prefix(p, cfg, obs, bf16, softmax_dtype)
    pre_attn = make_attn_mask(pre_mask, pre_ar)
    positions = torch.cumsum(pre_mask.to(torch.int64), 1) - 1
    cfgs = [cfg.gemma, cfg.expert]
    _, cache = gemma_forward(p, cfgs, [pre_tok, None], positions, pre_attn, [None, None], bf16)
    x_t = noise.clone()
    t = 1.0
    n_iter = 0
    while t >= -dt / 2:
        suf_tok, suf_mask, suf_ar, cond = embed_suffix(p, cfg, x_t, torch.full((B,), t, dtype=torch.float32))
        suf_attn = make_attn_mask(suf_mask, suf_ar[None, :])
        pre_part = pre_mask[:, None, :].expand(-1, suf_tok.shape[1], -1)
        full = torch.cat([pre_part, suf_attn], -1)
        pos = pre_mask.sum(-1)[:, None] + torch.cumsum(suf_mask.to(torch.int64), -1) - 1
        (_, suf_out), _ = gemma_forward(p, cfgs, [None, suf_tok], pos, full, [None, cond], bf16, kv_cache=cache)
        v_t = suf_out[:, -cfg.action_horizon:] @ p["action_out_proj/kernel"] + p["action_out_proj/bias"]
        x_t = x_t + dt * v_t
        t = t + dt
        n_iter += 1
    assert n_iter == num_steps
    return x_t


def left_to_right_align(x, input_mask, attn_mask):
    """OP/models/pi0_fast.py:52-64, per example: roll so that the last valid token sits in the last slot."""
    n = input_mask.shape[0]
    seqlen = int((input_mask.to(torch.int64) * torch.arange(n)).max()) + 1
    return (torch.roll(x, -seqlen, 0), torch.roll(input_mask, -seqlen, 0), torch.roll(attn_mask, (-seqlen, -seqlen), (0, 1)))


def sample_tokens(p, cfg, obs, *, max_decoding_steps=390, bf16: bool, softmax_dtype="bf16", return_logits=False,
                  temperature: float = 0.0, gumbel=None):
    """LAP.sample_tokens (lap.py:678-766): right-aligned prefix prefill -> KV cache -> one token per step through expert 0
    alone.  The decode-step mask is the reference's RANGE mask (slot >= prefix_start and slot <= current), not a validity
    mask (lap.py:737-741).  temperature = 0: greedy.  temperature > 0 (lap.py:727-729): jax.random.categorical(key, z) is
    argmax(z + Gumbel noise drawn from key); the noise is passed in explicitly (`gumbel` [B, S, V]) because the threefry
    stream cannot be reproduced."""
    pre_tok, pre_mask, pre_ar = embed_prefix(p, cfg, obs, bf16, softmax_dtype)
    attn = make_attn_mask(pre_mask, pre_ar)
    B, P, _ = pre_tok.shape
    al = [left_to_right_align(pre_tok[b], pre_mask[b], attn[b]) for b in range(B)]
    pre_tok, pre_mask, attn = (torch.stack([a[i] for a in al]) for i in range(3))
    prefill_len = pre_mask.sum(-1)
    prefix_start = P - prefill_len
    S = max_decoding_steps
    attn = torch.cat([attn, torch.zeros(B, P, S, dtype=torch.bool)], -1)
    positions = torch.cumsum(pre_mask.to(torch.int64), -1) - 1
    cfgs = [cfg.gemma, cfg.expert]
    # prefill: the cache holds P slots now, S more are appended one per step (keys beyond the current slot are masked)
    (pre_out, _), cache = gemma_forward(p, cfgs, [pre_tok, None], positions, attn[:, :, :P], [None, None], bf16)
    last_logit = decode_logits(p, pre_out[:, -1:], bf16)
    out = torch.zeros(B, S, dtype=torch.int64)
    eos = torch.zeros(B, dtype=torch.bool)
    logits_log = []
    step = 0
    while (not bool(eos.all())) and step < S:
        if temperature > 0.0:
            token = (last_logit / temperature + gumbel[:, step : step + 1].to(last_logit.dtype)).argmax(-1)
        else:
            token = last_logit.argmax(-1)  # [B, 1]
        logits_log.append(last_logit[:, 0])
        out[:, step] = token[:, 0]
        eos = eos | (token[:, 0] == 1)  # EOS_TOKEN (lap.py:32)
        emb = embed_tokens(p, cfg, token, bf16)
        pos = prefill_len[:, None] + step
        slots = torch.arange(P + step + 1)[None, None, :]
        mask = (slots >= prefix_start[:, None, None]) & (slots < P + step + 1)
        (o, _), new = gemma_forward(p, cfgs, [emb, None], pos, mask, [None, None], bf16, kv_cache=cache)
        cache = new  # gemma_attention returns (cache ++ new token)
        last_logit = decode_logits(p, o, bf16)
        step += 1
    return (out, torch.stack(logits_log, 1)) if return_logits else out


# ---------------------------------------------------------------------------------------------
# train step  (scripts/train.py:329-419; optax semantics restated, Appendix A.9)
# ---------------------------------------------------------------------------------------------
def is_kernel_param(name: str, t: torch.Tensor) -> bool:
    """scripts/train.py:402-409: Param, ndim>1, not .*/(bias|scale|pos_embedding|input_embedding)."""
    leaf = name.rsplit("/", 1)[-1]
    return t.ndim > 1 and leaf not in ("bias", "scale", "pos_embedding", "input_embedding")


def train_step(train_cfg, state: dict, obs, actions, noise, time, *, bf16: bool, softmax_dtype="bf16"):
    """One optimisation step.  state = {step:int, params:{}, mu:{}, nu:{}, ema:{}|None}.  Returns (state, info)."""
    cfg = train_cfg.model
    params = {k: v.detach().clone().requires_grad_(True) for k, v in state["params"].items()}
    loss, metrics = compute_loss(params, cfg, obs, actions, noise, time, bf16=bf16, softmax_dtype=softmax_dtype)
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in params.items()}
    gnorm = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).float()
    o = train_cfg.optimizer
    # optax.clip_by_global_norm
    clip = o.clip_gradient_norm
    scale = 1.0 if gnorm < clip else float(clip / gnorm)
    step = state["step"]
    lr = train_cfg.lr_schedule.lr(step)
    count = step + 1
    bc1, bc2 = 1.0 - o.b1 ** count, 1.0 - o.b2 ** count
    decay, ema_on = train_cfg.get_ema_decay_for_step(step)
    new = {"step": step + 1, "params": {}, "mu": {}, "nu": {}, "ema": None if state.get("ema") is None else {}}
    for k, pv in state["params"].items():
        g = grads[k] * scale
        mu = o.b1 * state["mu"][k] + (1 - o.b1) * g
        nu = o.b2 * state["nu"][k] + (1 - o.b2) * g * g
        upd = (mu / bc1) / (torch.sqrt(nu / bc2) + o.eps) + o.weight_decay * pv
        npv = pv - lr * upd
        new["params"][k], new["mu"][k], new["nu"][k] = npv, mu, nu
        if new["ema"] is not None:
            new["ema"][k] = decay * state["ema"][k] + (1 - decay) * npv if ema_on else state["ema"][k]
    pnorm = torch.sqrt(sum((v.double() ** 2).sum() for k, v in new["params"].items() if is_kernel_param(k, v))).float()
    info = {"loss": loss.detach(), "grad_norm": gnorm, "grad_norm_f32": gnorm, "param_norm": pnorm,
            **{k: v.detach() for k, v in metrics.items()}}
    return new, info, grads
