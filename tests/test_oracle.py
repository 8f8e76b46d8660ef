"""CPU tests of the oracle (test infrastructure) — known answers lifted from the reference + self-made pins."""
import math
import os

import numpy as np
import pytest
import torch

from lap_b200 import params as P
from lap_b200.config import CosineDecaySchedule, EmaScheduleChoice, TrainConfig, get_config
from lap_b200.data import synthetic_batch
from oracle import lap_oracle as O
from tests.helpers import obs_for_oracle, rel_err

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _t(x):
    return torch.tensor(x)


# ---- KAT (1): the three worked examples in make_attn_mask's docstring, OP/models/pi0.py:26-33 ----
def test_make_attn_mask_docstring_causal():
    n = 6
    m = O.make_attn_mask(torch.ones(1, n, dtype=torch.bool), _t([[1, 1, 1, 1, 1, 1]]).bool())
    assert torch.equal(m[0], torch.tril(torch.ones(n, n, dtype=torch.bool)))


def test_make_attn_mask_docstring_prefix_lm():
    m = O.make_attn_mask(torch.ones(1, 6, dtype=torch.bool), _t([[0, 0, 0, 1, 1, 1]]).bool())[0]
    exp = torch.tril(torch.ones(6, 6, dtype=torch.bool))
    exp[:3, :3] = True  # first 3 tokens attend among themselves, last 3 causal
    assert torch.equal(m, exp)
    # "The first entry could also be a 1 without changing behaviour"
    m2 = O.make_attn_mask(torch.ones(1, 6, dtype=torch.bool), _t([[1, 0, 0, 1, 1, 1]]).bool())[0]
    assert torch.equal(m, m2)


def test_make_attn_mask_docstring_blocks():
    ar = [1, 0, 1, 0, 1, 0, 0, 1, 0, 0]
    m = O.make_attn_mask(torch.ones(1, 10, dtype=torch.bool), _t([ar]).bool())[0]
    block = np.cumsum(ar)  # 4 causal blocks
    exp = torch.from_numpy(block[None, :] <= block[:, None])
    assert torch.equal(m, exp)
    assert len(set(block)) == 4


def test_make_attn_mask_padding():
    im = _t([[1, 1, 0, 1]]).bool()
    m = O.make_attn_mask(im, _t([[0, 0, 0, 0]]).bool())[0]
    assert not m[2].any() and not m[:, 2].any() and m[0, 3] and m[3, 0]


def _brute_combined(pm, par, la, A):
    """Appendix C semantics written as explicit loops."""
    P_ = len(pm)
    T = P_ + A
    cum = np.cumsum(par)
    M = np.zeros((T, T), bool)
    for i in range(P_):
        for j in range(P_):
            M[i, j] = pm[i] and pm[j] and cum[j] <= cum[i]
    for i in range(P_, T):
        for j in range(P_):
            M[i, j] = pm[j] and not la[j]
        for j in range(P_, T):
            M[i, j] = True
    pos = np.concatenate([np.cumsum(pm) - 1, (pm & ~la).sum() + np.arange(A)])
    return M, pos


def test_combined_mask_matches_appendix_c():
    rng = np.random.default_rng(0)
    for _ in range(5):
        n_img, L, A = 6, 12, 4
        pm = np.concatenate([np.repeat(rng.random() < 0.8, n_img), np.arange(L) < rng.integers(4, L)])
        la_text = np.zeros(L, bool)
        n_p = rng.integers(1, 4)
        la_text[n_p:n_p + rng.integers(1, 5)] = True
        la_text &= pm[n_img:]
        la = np.concatenate([np.zeros(n_img, bool), la_text])
        pmt, part, lat = _t(pm[None]), _t(la[None]), _t(la_text[None])
        pma = O.build_prefix_action_mask(pmt, lat)
        sm = torch.ones(1, A, dtype=torch.bool)
        sar = _t([[True] + [False] * (A - 1)])
        M = O.build_combined_attention_mask(pmt, part, pma, sm, sar)[0].numpy()
        pos = O.build_combined_positions(pmt, pma, sm)[0].numpy()
        Mb, posb = _brute_combined(pm, la, la, A)
        assert np.array_equal(M, Mb)
        assert np.array_equal(pos, posb)


def test_posemb_sincos_known_values():
    e = O.posemb_sincos(torch.tensor([0.0, 1.0]), 8, 4e-3, 4.0)
    assert e.shape == (2, 8)
    assert torch.allclose(e[0], torch.tensor([0, 0, 0, 0, 1, 1, 1, 1.0]))
    # last period is max_period=4 -> angle 2*pi/4
    assert abs(e[1, 3].item() - math.sin(2 * math.pi / 4.0)) < 1e-6


def test_rope_is_rotation_and_position_zero_identity():
    x = torch.randn(2, 5, 3, 16)
    pos = torch.zeros(2, 5, dtype=torch.int32)
    assert torch.allclose(O.apply_rope(x, pos, False), x)
    pos = torch.arange(5)[None].repeat(2, 1)
    y = O.apply_rope(x, pos, False)
    assert torch.allclose(y.norm(dim=-1), x.norm(dim=-1), atol=1e-5)


def test_lr_schedule_matches_optax_warmup_cosine():
    s = CosineDecaySchedule(warmup_steps=1000, peak_lr=5e-5, decay_steps=40_000, decay_lr=5e-5)
    assert abs(s.lr(0) - 5e-5 / 1001) < 1e-12
    assert abs(s.lr(1000) - 5e-5) < 1e-12 and abs(s.lr(39_999) - 5e-5) < 1e-12
    s2 = CosineDecaySchedule(warmup_steps=10, peak_lr=1.0, decay_steps=110, decay_lr=0.1)
    assert abs(s2.lr(5) - (1 / 11 + (1 - 1 / 11) * 0.5)) < 1e-9
    assert abs(s2.lr(60) - (0.1 + 0.9 * 0.5)) < 1e-9  # halfway through the cosine
    assert abs(s2.lr(10_000) - 0.1) < 1e-9


def test_ema_schedule():
    tc = TrainConfig(ema_decay=0.999, ema_schedule_choice=EmaScheduleChoice(kind="constant"))
    assert tc.get_ema_init() == (0.999, True) and tc.get_ema_decay_for_step(0) == (0.999, True)
    tc = TrainConfig(ema_decay=0.999, num_train_steps=100, ema_schedule_choice=EmaScheduleChoice(kind="cosine_delayed", start_step=50))
    assert tc.get_ema_init() == (0.0, True)
    d, on = tc.get_ema_decay_for_step(49)
    assert not on
    d, on = tc.get_ema_decay_for_step(75)
    assert on and abs(d - 0.999 * 0.5) < 1e-9


@pytest.fixture(scope="module")
def tiny():
    tc = get_config("debug_tiny")
    ref = P.init_reference_params(tc.model, 7, reference_zero_init=False)
    b = synthetic_batch(tc.model, 3, step=11)
    return tc, ref, b


def test_golden_reproduced(tiny):
    tc, ref, b = tiny
    g = np.load(os.path.join(GOLDEN, "debug_tiny_B3.npz"))
    t = lambda x: torch.from_numpy(np.asarray(x))
    for bf, tag in ((True, "bf16"), (False, "f32")):
        loss, m, aux = O.compute_loss(ref, tc.model, obs_for_oracle(b), t(b["actions"]), t(b["noise"]), t(b["time"]), bf16=bf, return_aux=True)
        assert abs(loss.item() - float(g[f"loss_{tag}"])) < 1e-5
        assert rel_err(aux["v_t"], g[f"v_t_{tag}"]) < 1e-4
    assert np.array_equal(np.packbits(aux["mask"].numpy(), axis=-1), g["mask"])
    assert np.array_equal(aux["positions"].numpy(), g["positions"])


def test_shapes_contract(tiny):
    """OP/models/model_test.py:20-24: sample_actions -> (B, ah, ad); loss is a scalar for LAP."""
    tc, ref, b = tiny
    t = lambda x: torch.from_numpy(np.asarray(x))
    a = O.sample_actions(ref, tc.model, obs_for_oracle(b, langact=False), t(b["noise"]), num_steps=10, bf16=False)
    assert a.shape == (3, tc.model.action_horizon, tc.model.action_dim)
    a5 = O.sample_actions(ref, tc.model, obs_for_oracle(b, langact=False), t(b["noise"]), num_steps=5, bf16=False)
    assert a5.shape == a.shape  # Euler loop ran exactly num_steps iterations (asserted inside)


def test_bf16_and_fp32_oracle_agree(tiny):
    tc, ref, b = tiny
    t = lambda x: torch.from_numpy(np.asarray(x))
    l32, _ = O.compute_loss(ref, tc.model, obs_for_oracle(b), t(b["actions"]), t(b["noise"]), t(b["time"]), bf16=False)
    l16, _ = O.compute_loss(ref, tc.model, obs_for_oracle(b), t(b["actions"]), t(b["noise"]), t(b["time"]), bf16=True)
    assert abs(l32.item() - l16.item()) < 5e-3 * abs(l32.item())


def test_stop_action_to_vlm_grad_blocks_exactly_the_action_loss_into_the_vlm(tiny):
    """gemma.py:206-213,242-269 (`lap` config, config.py:616): action-expert queries read expert 0's K/V through
    stop_gradient.  Consequences checked on the oracle: the forward value is unchanged; the action loss has ZERO gradient
    w.r.t. every VLM parameter (SigLIP, Gemma-2B, embedding); the action expert's own gradients and the language-loss
    gradients are what they were."""
    import dataclasses
    tc, ref, b = tiny
    t = lambda x: torch.from_numpy(np.asarray(x))
    args = (obs_for_oracle(b), t(b["actions"]), t(b["noise"]), t(b["time"]))

    def grads(cfg):
        p = {k: v.detach().clone().requires_grad_(True) for k, v in ref.items()}
        loss, _ = O.compute_loss(p, cfg, *args, bf16=False)
        loss.backward()
        return loss.detach(), {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in p.items()}

    base = dataclasses.replace(tc.model, language_loss_weight=0.0)  # action loss only
    l0, g0 = grads(base)
    l1, g1 = grads(dataclasses.replace(base, stop_action_to_vlm_grad=True))
    assert abs(l0.item() - l1.item()) < 1e-6 * abs(l0.item())
    expert = lambda k: (k.startswith("PaliGemma/llm/") and any(s in k for s in ("einsum_1", "mlp_1", "norm_1"))) or \
        k.split("/")[0] in ("action_in_proj", "action_out_proj", "time_mlp_in", "time_mlp_out")
    n_vlm = 0
    for k in ref:
        if expert(k):
            assert rel_err(g1[k], g0[k]) < 1e-5, k
        else:
            n_vlm += 1
            assert g1[k].abs().max() == 0, k  # nothing of the action loss reaches the VLM
    assert n_vlm > 10 and any(g0[k].abs().max() > 0 for k in ref if not expert(k))
    # language loss only: the flag changes nothing
    lang = dataclasses.replace(tc.model, action_loss_weight=0.0)
    _, ga = grads(lang)
    _, gb = grads(dataclasses.replace(lang, stop_action_to_vlm_grad=True))
    for k in ref:
        if ga[k].abs().max() > 0:
            assert rel_err(gb[k], ga[k]) < 1e-5, k


def test_training_path_equals_cached_inference_path(tiny):
    """Reference-implied invariant (SURVEY §8c ii): with no lang-action tokens, the suffix outputs of the joint
    [prefix,suffix] pass equal those of prefix-KV-cache + suffix-only pass."""
    tc, ref, b = tiny
    cfg = tc.model
    t = lambda x: torch.from_numpy(np.asarray(x))
    obs = obs_for_oracle(b, langact=False)
    B = 3
    x_t, time = t(b["noise"]), t(b["time"])
    suf_tok, suf_mask, suf_ar, cond = O.embed_suffix(ref, cfg, x_t, time)
    pre_tok, pre_mask, pre_ar = O.embed_prefix(ref, cfg, obs, False)
    cfgs = [cfg.gemma, cfg.expert]
    mask = O.build_combined_attention_mask(pre_mask, pre_ar, pre_mask, suf_mask, suf_ar[None].expand(B, -1))
    pos = O.build_combined_positions(pre_mask, pre_mask, suf_mask)
    (_, joint), _ = O.gemma_forward(ref, cfgs, [pre_tok, suf_tok], pos, mask, [None, cond], False)
    pattn = O.make_attn_mask(pre_mask, pre_ar)
    ppos = torch.cumsum(pre_mask.long(), 1) - 1
    _, cache = O.gemma_forward(ref, cfgs, [pre_tok, None], ppos, pattn, [None, None], False)
    full = torch.cat([pre_mask[:, None, :].expand(-1, suf_tok.shape[1], -1), O.make_attn_mask(suf_mask, suf_ar[None])], -1)
    spos = pre_mask.sum(-1)[:, None] + torch.cumsum(suf_mask.long(), -1) - 1
    (_, cached), _ = O.gemma_forward(ref, cfgs, [None, suf_tok], spos, full, [None, cond], False, kv_cache=cache)
    assert rel_err(cached, joint) < 1e-5


def test_padding_invariance(tiny):
    """Outputs at valid positions do not depend on token ids / pixels at masked positions (SURVEY §8c iv)."""
    tc, ref, b = tiny
    t = lambda x: torch.from_numpy(np.asarray(x))
    l0, _ = O.compute_loss(ref, tc.model, obs_for_oracle(b), t(b["actions"]), t(b["noise"]), t(b["time"]), bf16=False)
    b2 = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in b.items()}
    pad = ~b2["tokenized_prompt_mask"]
    b2["tokenized_prompt"] = np.where(pad, 5, b2["tokenized_prompt"]).astype(np.int32)
    l1, _ = O.compute_loss(ref, tc.model, obs_for_oracle(b2), t(b["actions"]), t(b["noise"]), t(b["time"]), bf16=False)
    assert abs(l0.item() - l1.item()) < 1e-5


def test_adamw_ema_closed_form():
    """One optimizer step against hand-computed optax semantics."""
    tc = get_config("debug_tiny")
    ref = P.init_reference_params(tc.model, 1, reference_zero_init=False)
    b = synthetic_batch(tc.model, 2, step=0)
    t = lambda x: torch.from_numpy(np.asarray(x))
    z = {k: torch.zeros_like(v) for k, v in ref.items()}
    state = dict(step=0, params=ref, mu=z, nu=dict(z), ema={k: v.clone() for k, v in ref.items()})
    ns, info, g = O.train_step(tc, state, obs_for_oracle(b), t(b["actions"]), t(b["noise"]), t(b["time"]), bf16=False)
    k = "action_out_proj/kernel"
    o = tc.optimizer
    gn = info["grad_norm"].item()
    gg = g[k] * (1.0 if gn < 1.0 else 1.0 / gn)
    mu, nu = 0.1 * gg, 0.05 * gg * gg
    upd = (mu / 0.1) / (torch.sqrt(nu / 0.05) + o.eps) + o.weight_decay * ref[k]
    exp = ref[k] - tc.lr_schedule.lr(0) * upd
    assert torch.allclose(ns["params"][k], exp, rtol=1e-5, atol=1e-8)
    assert torch.allclose(ns["ema"][k], 0.999 * ref[k] + 0.001 * exp, rtol=1e-6, atol=1e-8)


def test_train_step_matches_an_independent_adamw(tiny):
    """a3 pin against an independent implementation (optax itself is not installable): two oracle train steps vs
    torch.optim.AdamW (decoupled weight decay, the same update optax.adamw applies: p - lr*(m_hat/(sqrt(v_hat)+eps) + wd*p))
    fed with the same clipped gradients, learning rate from the warm-up schedule."""
    tc, ref, b = tiny
    t = lambda x: torch.from_numpy(np.asarray(x))
    args = (obs_for_oracle(b), t(b["actions"]), t(b["noise"]), t(b["time"]))
    z = {k: torch.zeros_like(v) for k, v in ref.items()}
    state = dict(step=0, params={k: v.clone() for k, v in ref.items()}, mu=z, nu={k: v.clone() for k, v in z.items()}, ema=None)
    tp = {k: v.clone().requires_grad_(True) for k, v in ref.items()}
    o = tc.optimizer
    opt = torch.optim.AdamW(list(tp.values()), lr=1.0, betas=(o.b1, o.b2), eps=o.eps, weight_decay=o.weight_decay)
    for step in range(2):
        state, info, grads = O.train_step(tc, state, *args, bf16=False)
        gn = float(info["grad_norm"])
        scale = 1.0 if gn < o.clip_gradient_norm else o.clip_gradient_norm / gn  # optax.clip_by_global_norm
        for k, v in tp.items():
            v.grad = grads[k] * scale
        for grp in opt.param_groups:
            grp["lr"] = tc.lr_schedule.lr(step)
        opt.step()
        for k in ref:
            d_o, d_t = state["params"][k] - ref[k], tp[k].detach() - ref[k]
            # (the difference p_new - p cancels ~4 digits of fp32 for parameters of magnitude 1)
            assert rel_err(d_o, d_t) < 1e-3, (step, k, rel_err(d_o, d_t))


def test_sample_tokens_equals_teacher_forced_forward(tiny):
    """AR decode (lap.py:678-766) invariant: the logits of greedy decode step s equal the training-style forward of the same
    prompt with the generated tokens appended as causal lang-action tokens (prefix-LM mask, positions cumsum-1) — checks
    the right-aligned prefill, the KV cache growth, the range mask and the positions of the decode loop."""
    tc, ref, _ = tiny
    cfg = tc.model
    b = synthetic_batch(cfg, 3, step=5, with_langact=False)
    L = cfg.max_token_len
    n_p = np.array([7, 10, 12])
    b["tokenized_prompt_mask"] = np.arange(L)[None, :] < n_p[:, None]
    b["image_mask"] = {k: np.ones_like(v) for k, v in b["image_mask"].items()}  # no holes: range mask == validity mask
    t = lambda x: torch.from_numpy(np.asarray(x))
    K = 6
    obs = obs_for_oracle(b, langact=False)
    toks, logits = O.sample_tokens(ref, cfg, obs, max_decoding_steps=K, bf16=False, return_logits=True)
    assert toks.shape == (3, K) and logits.shape[:2] == (3, K)
    # teacher forcing
    prompt = b["tokenized_prompt"].copy()
    pm, la = b["tokenized_prompt_mask"].copy(), np.zeros((3, L), dtype=bool)
    for i in range(3):
        prompt[i, n_p[i]: n_p[i] + K] = toks[i].numpy()
        pm[i, n_p[i]: n_p[i] + K] = True
        la[i, n_p[i]: n_p[i] + K] = True
    b2 = dict(b, tokenized_prompt=prompt, tokenized_prompt_mask=pm, tokenized_langact_mask=la)
    _, _, aux = O.compute_loss(ref, cfg, obs_for_oracle(b2), t(b["actions"]), t(b["noise"]), t(b["time"]), bf16=False,
                               return_aux=True)
    for i in range(3):
        for s_ in range(K):
            assert rel_err(logits[i, s_], aux["logits"][i, n_p[i] + s_ - 1]) < 1e-4, (i, s_)
    # a dropped camera leaves a hole inside the right-aligned range: the decode still runs and returns tokens
    b["image_mask"]["left_wrist_0_rgb"][1] = False
    toks2 = O.sample_tokens(ref, cfg, obs_for_oracle(b, langact=False), max_decoding_steps=3, bf16=True)
    assert toks2.shape == (3, 3)


def test_sample_tokens_temperature_is_gumbel_argmax(tiny):
    """lap.py:727-729: jax.random.categorical(key, z / T) == argmax(z / T + Gumbel noise).  With zero noise the sampled
    decode is the greedy decode; with huge noise on one token per step it emits exactly those tokens; the teacher-forcing
    invariant keeps holding for whatever was sampled (the sampled token, not the argmax, is fed back)."""
    tc, ref, _ = tiny
    cfg = tc.model
    b = synthetic_batch(cfg, 2, step=5, with_langact=False)
    obs = obs_for_oracle(b, langact=False)
    K = 4
    greedy = O.sample_tokens(ref, cfg, obs, max_decoding_steps=K, bf16=False)
    zero = torch.zeros(2, K, cfg.vocab_size)
    assert torch.equal(O.sample_tokens(ref, cfg, obs, max_decoding_steps=K, bf16=False, temperature=0.5, gumbel=zero), greedy)
    forced = torch.tensor([[5, 9, 11, 3], [7, 7, 2, 4]])
    g = torch.zeros(2, K, cfg.vocab_size)
    g.scatter_(2, forced[:, :, None], 1e9)
    toks, logits = O.sample_tokens(ref, cfg, obs, max_decoding_steps=K, bf16=False, temperature=1.3, gumbel=g, return_logits=True)
    assert torch.equal(toks, forced)
    toks2, logits2 = O.sample_tokens(ref, cfg, obs, max_decoding_steps=K, bf16=False, return_logits=True)
    assert rel_err(logits[:, 0], logits2[:, 0]) < 1e-6           # the first logits do not depend on what is sampled
    assert rel_err(logits[:, 1], logits2[:, 1]) > 1e-3           # later ones do


def test_image_augmentation_statement_properties():
    """oracle/image_oracle.py (model_adapter.py:118-151 with explicit parameters): a skipped (VQA) sample is untouched; a
    constant image stays constant under the geometric part (away from the zero-filled corners) and moves by exactly the
    brightness / contrast formula; the output stays in [-1, 1]; a rotation by +a then reading the centre pixel is a no-op."""
    from oracle import image_oracle as IO
    rng = np.random.default_rng(0)
    img = rng.uniform(-1, 1, (3, 32, 32, 3)).astype(np.float32)
    p = IO.draw_params(rng, 3, 32, 32, skip=[0, 1, 0])
    out = IO.augment(img, p)
    assert out.shape == img.shape and np.array_equal(out[1], img[1]) and np.abs(out).max() <= 1.0
    const = np.full((1, 32, 32, 3), 0.2, np.float32)
    q = np.array([[0.5, 0.7, 3.0, 0.1, -0.1, 0.2, 0]], np.float32)
    v = 0.2 * 0.5 + 0.5
    v = v * (1 - 0.1) + 0.1
    v = (v - 0.5) * (1 - 0.1) + 0.5
    assert np.allclose(IO.augment(const, q)[0, 8:24, 8:24], v * 2 - 1, atol=1e-6)   # grey: saturation is a no-op
    z = np.array([[0.0, 0.0, 0.0, 0, 0, 0, 0]], np.float32)                        # crop at (0, 0), no rotation / jitter
    zo = IO.augment(img[:1], z)[0]
    # pure zoom: output pixel (y, x) reads ((y + .5) * 30/32 - .5, ...): pixel (0, 0) reads (-1/32, -1/32) -> corner weight
    w = 1 - 1 / 32
    assert np.allclose(zo[0, 0], (img[0, 0, 0] * 0.5 + 0.5) * w * w * 2 - 1, atol=1e-6)


def test_resize_plan_is_the_sparse_form_of_the_resize_matrices():
    from lap_b200 import image_tools as it
    for (h, w) in ((480, 640), (100, 150), (224, 300)):
        plan = it.resize_plan(h, w, 224, 224)
        for dense, start, vals, taps in ((it._weights(h, plan["rh"], True), plan["ystart"], plan["yw"], plan["ytaps"]),
                                        (it._weights(w, plan["rw"], True), plan["xstart"], plan["xw"], plan["xtaps"])):
            rec = np.zeros_like(dense)
            for p in range(dense.shape[0]):
                for t in range(taps):
                    if start[p] + t < dense.shape[1]:
                        rec[p, start[p] + t] += vals[p, t]
            assert np.array_equal(rec, dense)
        assert plan["ph0"] * 2 + plan["rh"] in (224, 223) and plan["pw0"] * 2 + plan["rw"] in (224, 223)
