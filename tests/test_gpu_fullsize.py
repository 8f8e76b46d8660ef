"""Full-size parity on hardware: the engine at `get_config("lap_libero")` — the real LAP-3B (SigLIP-So400m + Gemma-2B +
Gemma-300M expert, vocabulary 257 152, prefix 692 tokens) — against the bf16-emulating oracle on the same seeded weights
and inputs.  Covers the shapes the small debug configs cannot: the LM head at N = 257 152 (hi/lo split table), the
M = 692 / K = 16 384 GEMMs inside the real network, 27 SigLIP and 18 Gemma layers of accumulated bf16 rounding, K1 at
T = 702 and K10 on the real expert.  Reference: src/lap/models/lap.py:380-675.

The oracle's forward at this size takes ~1 minute on the host cores, so everything hangs off ONE module-scoped fixture
(weights built once, oracle run once per quantity)."""
import dataclasses
import gc

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from lap_b200 import params as P  # noqa: E402
from lap_b200.config import get_config  # noqa: E402
from lap_b200.data import synthetic_batch  # noqa: E402
from oracle import lap_oracle as O  # noqa: E402
from tests.helpers import obs_for_oracle, rel_err  # noqa: E402

# north_star: "<= 1e-3 relative for bf16 logits and actions".  Scalars (losses, metrics) meet it.  Activations cannot:
# at full depth with random-init weights two CORRECT bf16 implementations differ by ~1e-2 normwise — the oracle against
# ITSELF with a different BLAS thread count gives 1.08e-2 on v_t (tools/parity_floor.py, profiles/r02_parity_floor.md).
# Activation-level asserts are therefore (i) an absolute bound of ~2x that floor against the bf16 oracle and (ii) the
# engine must be no further from the fp32 oracle (the mathematical function) than 1.5x the bf16 oracle itself is.
TOL_LOSS = 1e-3
TOL_VT = 2.5e-2
TOL_ACTIONS = 5e-3
REL_TO_FP32 = 1.5
_t = lambda x: torch.from_numpy(np.asarray(x))


def _report(name, value, tol):
    print(f"[fullsize-parity] {name}: {value:.3e} (tol {tol:.1e})")


@pytest.fixture(scope="module")
def full():
    from lap_b200.model import LAP

    tc = get_config("lap_libero")
    cfg = tc.model
    ref = P.init_reference_params(cfg, 7, reference_zero_init=False)
    model = LAP(cfg, init=False)
    model.load_params(ref)
    yield tc, cfg, ref, model
    del model, ref
    gc.collect()
    torch.cuda.empty_cache()


def test_full_size_loss_matches_oracle(full):
    """compute_loss on two samples (one with a dropped wrist camera, ragged prompt / lang-action lengths): total loss and
    the three metrics against O.compute_loss(bf16=True); token / patch indexing (mask, positions) bit-exact."""
    from lap_b200 import ops
    from lap_b200.train import batch_from_dict

    tc, cfg, ref, model = full
    B = 2
    b = synthetic_batch(cfg, B, step=11)
    b["sample_mask"][:] = True
    b["image_mask"]["left_wrist_0_rgb"][1] = False
    b["image"]["left_wrist_0_rgb"][1] = -1.0
    obs, actions, extra = batch_from_dict(b)
    loss, m = model.compute_loss(0, obs, actions, noise=extra["noise"], time=extra["time"])
    with torch.no_grad():
        loss_o, m_o, aux = O.compute_loss(ref, cfg, obs_for_oracle(b), _t(b["actions"]), _t(b["noise"]), _t(b["time"]),
                                          bf16=True, return_aux=True)
    e = abs(loss.item() - float(loss_o)) / abs(float(loss_o))
    _report("loss", e, TOL_LOSS)
    assert e < TOL_LOSS
    for k in ("lang_loss", "action_loss", "langact_loss"):
        e = abs(m[k].item() - float(m_o[k])) / abs(float(m_o[k]))
        _report(k, e, TOL_LOSS)
        assert e < TOL_LOSS, k
    T = cfg.prefix_len + cfg.action_horizon
    Tpad = (T + 63) // 64 * 64
    dense = torch.zeros(B, T, T, dtype=torch.uint8, device="cuda")
    ops.mask_expand(model._bufs["mask.bits"], dense, B * T, T, Tpad // 32)
    assert np.array_equal(dense.cpu().numpy().astype(bool), aux["mask"].numpy())
    assert np.array_equal(model._bufs["mask.pos"].cpu().numpy(), aux["positions"].numpy())
    v_eng = model._bufs["loss.v"].view(B, cfg.action_horizon, -1)
    e = rel_err(v_eng, aux["v_t"])
    _report("v_t (suffix velocity, bf16 activations) vs bf16 oracle", e, TOL_VT)
    assert e < TOL_VT
    with torch.no_grad():
        loss32, _, aux32 = O.compute_loss(ref, cfg, obs_for_oracle(b), _t(b["actions"]), _t(b["noise"]), _t(b["time"]),
                                          bf16=False, return_aux=True)
    e_eng, e_orc = rel_err(v_eng, aux32["v_t"]), rel_err(aux["v_t"], aux32["v_t"])
    _report("v_t vs fp32 oracle: engine", e_eng, REL_TO_FP32 * e_orc)
    _report("v_t vs fp32 oracle: bf16 oracle", e_orc, float("nan"))
    assert e_eng < REL_TO_FP32 * e_orc
    assert abs(loss.item() - float(loss32)) < 3 * TOL_LOSS * abs(float(loss32))


def test_full_size_sample_actions_matches_oracle(full):
    """sample_actions at batch 1 (the serving case: CUDA graph + K10 persistent denoise loop) against the oracle."""
    from lap_b200.observation import Observation

    tc, cfg, ref, model = full
    b = synthetic_batch(cfg, 1, step=5, with_langact=False)
    obs = Observation.from_dict(b)
    with torch.no_grad():
        a_o = O.sample_actions(ref, cfg, obs_for_oracle(b, langact=False), _t(b["noise"]), num_steps=10, bf16=True)
        a_32 = O.sample_actions(ref, cfg, obs_for_oracle(b, langact=False), _t(b["noise"]), num_steps=10, bf16=False)
    a1 = model.sample_actions(0, obs, num_steps=10, noise=b["noise"])  # eager
    a2 = model.sample_actions(0, obs, num_steps=10, noise=b["noise"])  # capture + replay
    a3 = model.sample_actions(0, obs, num_steps=10, noise=b["noise"])  # replay
    assert model.denoise_error_flag() == 0
    assert a1.shape == (1, cfg.action_horizon, cfg.action_dim) and torch.isfinite(a1).all()
    e = rel_err(a1, a_o)
    _report("sample_actions B=1 (K10 path) vs bf16 oracle", e, TOL_ACTIONS)
    assert e < TOL_ACTIONS
    e_eng, e_orc = rel_err(a1, a_32), rel_err(a_o, a_32)
    _report("sample_actions vs fp32 oracle: engine", e_eng, REL_TO_FP32 * e_orc)
    _report("sample_actions vs fp32 oracle: bf16 oracle", e_orc, float("nan"))
    assert e_eng < REL_TO_FP32 * e_orc
    assert torch.equal(a1, a2) and torch.equal(a2, a3)
    # the kernel-per-op denoise path (what batch > 1 uses) on the same inputs
    model.use_denoise_megakernel = False
    model.use_cuda_graph = False
    try:
        a4 = model.sample_actions(0, obs, num_steps=10, noise=b["noise"])
    finally:
        model.use_denoise_megakernel = True
        model.use_cuda_graph = True
    e = rel_err(a4, a_o)
    _report("sample_actions B=1 (per-op path)", e, TOL_ACTIONS)
    assert e < TOL_ACTIONS


def test_full_size_gradients_match_oracle(full):
    """The hand-written backward at the REAL shapes (one sample: M = 692 prefix rows, T = 702 attention, LM-head weight
    gradient over 257 152 rows, 27 + 18 layers): every parameter gradient against torch autograd through the bf16-emulating
    oracle, normwise per tensor, plus the global gradient norm.  (scripts/train.py:351-361: value_and_grad of compute_loss)"""
    from lap_b200.train import batch_from_dict

    tc, cfg, ref, model = full
    b = synthetic_batch(cfg, 1, step=21)
    b["sample_mask"][:] = True
    obs, actions, extra = batch_from_dict(b)
    model.G = torch.zeros(model.layout.total, dtype=torch.float32, device=model.device)
    st = model._stage(obs, actions, extra["noise"], extra["time"], with_loss=True)
    loss = model.forward_backward(st)
    torch.cuda.synchronize()
    g_eng = model.params_reference(model.G)
    model.G = None
    params = {k: v.clone().requires_grad_(True) for k, v in ref.items()}
    loss_o, _ = O.compute_loss(params, cfg, obs_for_oracle(b), _t(b["actions"]), _t(b["noise"]), _t(b["time"]), bf16=True)
    loss_o.backward()
    e = abs(float(loss[0]) - float(loss_o)) / abs(float(loss_o))
    _report("B=1 train-forward loss", e, TOL_LOSS)
    assert e < TOL_LOSS
    gnorm = float(torch.sqrt(sum((v.grad.double() ** 2).sum() for v in params.values() if v.grad is not None)))
    worst, n_cmp = 0.0, 0
    for k, v in params.items():
        g_o = v.grad if v.grad is not None else torch.zeros_like(v)
        assert torch.isfinite(g_eng[k]).all(), k
        if float(g_o.norm()) < 2e-3 * gnorm:      # e.g. SigLIP key bias (softmax is shift-invariant): compare absolutely
            assert float((g_eng[k] - g_o).norm()) < 4e-3 * gnorm, k
        else:
            r = rel_err(g_eng[k], g_o)
            worst = max(worst, r)
            n_cmp += 1
            assert r < 4e-2, (k, r)               # bf16 cotangents through 45 layers vs fp32 autograd (measured: 1.6e-2)
        v.grad = None
    tot = float(torch.sqrt(sum((v.double() ** 2).sum() for v in g_eng.values())))
    _report(f"worst per-tensor gradient error over {n_cmp} tensors", worst, 4e-2)
    _report("global gradient norm", abs(tot - gnorm) / gnorm, 1e-2)
    assert abs(tot - gnorm) < 1e-2 * gnorm


def test_full_size_bj_shape_matches_oracle(full):
    """BASELINE.json's 48-token / 50-step / action_dim-32 shape at FULL width (the upstream Pi0Config defaults,
    OP/models/pi0_config.py:25-37): same towers, different prefix / suffix lengths and action projections."""
    from lap_b200.model import LAP
    from lap_b200.observation import Observation
    from lap_b200.train import batch_from_dict

    tc, cfg, ref, model = full
    cfg_bj = dataclasses.replace(cfg, action_dim=32, action_horizon=50, max_token_len=48, language_loss_weight=1.0)
    ref_bj = dict(ref)
    gen = torch.Generator().manual_seed(3)
    for k, s in P.reference_shapes(cfg_bj).items():  # only the action in / out projections change shape
        if tuple(ref[k].shape) != tuple(s):
            assert k.startswith("action_"), k
            std = 0.02 if k.endswith("bias") else 1.0 / (s[0] ** 0.5)
            ref_bj[k] = torch.randn(s, generator=gen) * std
    # the full-size model above is not needed any more in this (last) test: give its 33 GB back first
    model.P = model.W16 = model.E_split = None
    model._pool.clear(); model._bufs.clear(); model._io.clear(); model._infer_graphs.clear()
    gc.collect(); torch.cuda.empty_cache()
    m_bj = LAP(cfg_bj, init=False)
    m_bj.load_params(ref_bj)
    b = synthetic_batch(cfg_bj, 1, step=2)
    b["sample_mask"][:] = True
    obs, actions, extra = batch_from_dict(b)
    loss, m = m_bj.compute_loss(0, obs, actions, noise=extra["noise"], time=extra["time"])
    with torch.no_grad():
        loss_o, m_o = O.compute_loss(ref_bj, cfg_bj, obs_for_oracle(b), _t(b["actions"]), _t(b["noise"]), _t(b["time"]),
                                     bf16=True)
        a_o = O.sample_actions(ref_bj, cfg_bj, obs_for_oracle(b, langact=False), _t(b["noise"]), num_steps=10, bf16=True)
    e = abs(loss.item() - float(loss_o)) / abs(float(loss_o))
    _report("BJ-shape loss", e, TOL_LOSS)
    assert e < TOL_LOSS
    b2 = {k: v for k, v in b.items() if k != "tokenized_langact_mask"}
    a = m_bj.sample_actions(0, Observation.from_dict(b2), num_steps=10, noise=b["noise"])
    assert a.shape == (1, 50, 32)
    e = rel_err(a, a_o)
    _report("BJ-shape sample_actions", e, TOL_ACTIONS)
    assert e < TOL_ACTIONS
