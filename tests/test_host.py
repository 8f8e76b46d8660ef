"""CPU tests of the host side: parameter tree mapping, layout, C-ABI library loading and symbol export, error paths."""
import ctypes
import os
import math

import numpy as np
import pytest
import torch

from lap_b200 import _lib
from lap_b200 import params as P
from lap_b200.config import LAPConfig, TrainConfig, get_config
from lap_b200.data import synthetic_batch
from lap_b200.observation import CoTObservation, Observation


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    syms = _lib.exported_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), s
    assert lib.lapb200_version() >= 100
    assert isinstance(_lib.last_error(), str)


def test_gemm_struct_matches_header_size():
    # 64-bit fields are 8-aligned; a mismatch with include/lapb200.h would shift every later field
    assert ctypes.sizeof(_lib.GemmParams) % 8 == 0
    names = [f[0] for f in _lib.GemmParams._fields_]
    import re
    hdr = (_lib.INCLUDE / "lapb200.h").read_text()
    body = hdr[hdr.index("typedef struct {"):hdr.index("} lapb_gemm_t;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    decl = []
    for stmt in body.split(";"):
        stmt = stmt.replace("typedef struct {", "").strip()
        if not stmt:
            continue
        parts = stmt.split(",")
        for i, p_ in enumerate(parts):
            decl.append(re.findall(r"[A-Za-z_0-9]+", p_)[-1])
    assert decl == names


def test_ops_reject_cpu_tensors():
    from lap_b200 import ops
    x = torch.zeros(16, dtype=torch.float32)
    with pytest.raises(RuntimeError):
        ops.cast_f32_bf16(x, torch.zeros(16, dtype=torch.bfloat16))


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_model_fails_loudly_without_gpu():
    from lap_b200.model import LAP
    with pytest.raises(RuntimeError):
        LAP(get_config("debug_tiny").model)


def test_reference_engine_roundtrip_is_exact():
    for name in ("debug_tiny", "debug_small"):
        cfg = get_config(name).model
        ref = P.init_reference_params(cfg, 3, reference_zero_init=False)
        eng = P.reference_to_engine(cfg, ref)
        lay = P.FlatLayout(cfg)
        for k, s in lay.shapes.items():
            assert tuple(eng[k].shape) == tuple(s), k
        flat = P.flat_from_engine(lay, eng)
        back = P.engine_to_reference(cfg, P.engine_from_flat(lay, flat))
        assert set(back) == set(ref)
        for k in ref:
            assert torch.equal(back[k], ref[k]), k
        for k, o in lay.offsets.items():
            assert o % P.ALIGN == 0
        assert lay.kernel_end <= lay.offsets["g.embed"] and lay.total % 4 == 0


def test_lap3b_parameter_count_and_names():
    """3.353 B parameters; names/shapes of the checkpoint contract."""
    cfg = get_config("lap_libero").model
    shapes = P.reference_shapes(cfg)
    n = sum(math.prod(s) for s in shapes.values())
    assert abs(n / 1e9 - 3.353) < 0.002
    assert shapes["PaliGemma/llm/layers/attn/q_einsum/w"] == (18, 8, 2048, 256)
    assert shapes["PaliGemma/llm/layers/attn/kv_einsum_1/w"] == (18, 2, 1, 1024, 256)
    assert shapes["PaliGemma/llm/layers/mlp/gating_einsum"] == (18, 2, 2048, 16384)
    assert shapes["PaliGemma/img/Transformer/encoderblock/MlpBlock_0/Dense_0/kernel"] == (27, 1152, 4304)
    assert shapes["PaliGemma/llm/layers/pre_ffw_norm_1/Dense_0/kernel"] == (18, 1024, 3072)
    assert "PaliGemma/llm/layers/pre_ffw_norm_1/scale" not in shapes
    assert shapes["PaliGemma/llm/embedder/input_embedding"] == (257152, 2048)
    assert cfg.prefix_len == 692 and cfg.num_patches == 256


def test_nested_tree_helpers():
    flat = {"a/b/c": torch.ones(1), "a/d": torch.zeros(2)}
    nested = P.to_nested(flat)
    assert set(nested["a"].keys()) == {"b", "d"}
    nested["a"]["d"] = {"value": nested["a"]["d"]}  # orbax-style `value` leaf
    back = P.from_nested(nested)
    assert set(back) == set(flat)


def test_observation_from_dict_and_synthetic_batch():
    cfg = get_config("debug_tiny").model
    b = synthetic_batch(cfg, 4, step=2)
    obs = CoTObservation.from_dict(b)
    assert obs.tokenized_prompt.shape == (4, cfg.max_token_len)
    assert set(obs.images) == set(cfg.image_keys)
    la, pm = b["tokenized_langact_mask"], b["tokenized_prompt_mask"]
    assert (la & ~pm).sum() == 0 and la.any(axis=1).all()
    with pytest.raises(ValueError):
        Observation.from_dict({"image": {}, "image_mask": {}, "state": None, "tokenized_prompt": 1})
    b8 = synthetic_batch(cfg, 2, uint8_images=True)
    assert b8["image"]["base_0_rgb"].dtype == np.uint8


def test_config_registry():
    c = get_config("lap_libero")
    assert c.model.action_horizon == 10 and c.model.max_token_len == 180 and c.model.language_loss_weight == 0.4
    assert c.optimizer.weight_decay == 1e-4 and c.ema_schedule_choice.kind == "constant"
    with pytest.raises(ValueError):
        get_config("nope")


def test_transforms_match_reference_source():
    """lap_b200.transforms (Normalize / Unnormalize / PadStates) against the reference classes executed from
    src/lap/transforms.py (fixture: tests/golden/reference_transforms.npz, make_reference_transforms_golden.py)."""
    import os
    from lap_b200 import transforms as T
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_transforms.npz"))
    stats = T.NormStats(**{k: g[f"stats/{k}"] for k in ("mean", "std", "q01", "q99", "min", "max")})
    ns = {k: stats for k in ("state", "actions", "wide")}
    data = {k: g[f"data/{k}"] for k in ("state", "actions", "wide", "untouched")}
    for kind in ("normal", "bounds", "bounds_q99"):
        out = T.Normalize({k: v for k, v in ns.items() if k != "wide"}, kind)({k: v.copy() for k, v in data.items()})
        for k in data:
            np.testing.assert_array_equal(np.asarray(out[k]), g[f"normalize/{kind}/{k}"])
        out = T.Unnormalize(ns, T.NormalizationType(kind))({k: v.copy() for k, v in data.items()})
        for k in data:
            np.testing.assert_array_equal(np.asarray(out[k]), g[f"unnormalize/{kind}/{k}"])
    np.testing.assert_array_equal(T.PadStates(32)({"state": data["state"].copy()})["state"], g["padstates/short"])
    np.testing.assert_array_equal(T.PadStates(4)({"state": data["state"].copy()})["state"], g["padstates/long"])
    # round trip and error behaviour
    x = T.Unnormalize(ns, "bounds_q99")(T.Normalize(ns, "bounds_q99")({"actions": data["actions"].copy()}))["actions"]
    keep = stats.q01 != stats.q99
    np.testing.assert_allclose(x[:, keep], data["actions"][:, keep], rtol=1e-9, atol=1e-9)
    with pytest.raises(ValueError):
        T.Normalize({"state": T.NormStats(mean=stats.mean, std=stats.std)}, "bounds_q99")
    with pytest.raises(ValueError):
        T.Normalize(ns, "normal", strict=True)({"state": data["state"]})
    assert T.Normalize(None)({"a": 1}) == {"a": 1}


def test_tokenizer_masks_match_reference_source():
    """N2 second slice: `CoTTokenizer.tokenize` against `PaligemmaTokenizer.tokenize` of the reference executed from source
    with its real "lap" prompt format (tests/golden/make_reference_tokenizer_golden.py).  Bit-exact: tokens and all five masks,
    incl. truncation inside the prompt / inside the reasoning and the seeded reasoning dropout."""
    sentencepiece = pytest.importorskip("sentencepiece")
    from lap_b200 import tokenizer as tk
    from tests.golden.make_reference_tokenizer_golden import CASES

    gold = os.path.join(os.path.dirname(__file__), "golden")
    z = np.load(os.path.join(gold, "reference_tokenizer.npz"))
    sp = sentencepiece.SentencePieceProcessor(model_file=os.path.join(gold, "tiny_sp.model"))
    assert int(z["n_cases"]) == len(CASES)

    class StoredFormat:  # the prompt STRING is the injected part; everything after it is under test
        def __init__(self, text):
            self.text = text
        def format_prompt(self, prompt, state, state_type, **kw):
            return self.text
        direction_token_checker = staticmethod(tk.is_direction_natural)

    seen_number = seen_direction = seen_drop = 0
    for i, c in enumerate(CASES):
        fmt = StoredFormat(bytes(z[f"{i}/formatted"]).decode())
        t = tk.CoTTokenizer(sp, max_len=c["max_len"], prompt_format=fmt, reasoning_mask_prob=c.get("reasoning_mask_prob", 0.0))
        if "seed" in c:
            np.random.seed(c["seed"])
        res = t.tokenize(c["prompt"], c["reasoning"], state=c.get("state"), state_type=c.get("state_type"))
        # ... and the whole chain with this repo's own "lap" format instead of the stored string
        t2 = tk.CoTTokenizer(sp, max_len=c["max_len"], prompt_format="lap", reasoning_mask_prob=c.get("reasoning_mask_prob", 0.0))
        if "seed" in c:
            np.random.seed(c["seed"])
        res2 = t2.tokenize(c["prompt"], c["reasoning"], state=c.get("state"), state_type=c.get("state_type"))
        assert all((a is None and b is None) or np.array_equal(a, b) for a, b in zip(res, res2))
        for name, v in zip(("tokens", "attn", "reasoning", "number", "direction", "loss"), res):
            key = f"{i}/{name}"
            if v is None:
                assert key not in z, key
                continue
            assert v.dtype == z[key].dtype and v.shape == z[key].shape, key
            np.testing.assert_array_equal(v, z[key], err_msg=key)
        assert res[0].dtype == np.int32 and len(res[0]) == c["max_len"]
        if res[3] is not None:
            seen_number += int(res[3].sum()); seen_direction += int(res[4].sum())
            seen_drop += int((~res[5]).sum())
        expect = [p for p in ("right", "-", "+3", "back", "7", "cm") if tk.is_direction_natural(p)]
        assert "\x00".join(expect).encode() == bytes(z[f"{i}/direction_pieces"])
    assert seen_number > 0 and seen_direction > 0 and seen_drop > 0   # the fixture exercises every mask
    # decode() drops out-of-vocabulary ids (tokenizer.py:317-326)
    t = tk.CoTTokenizer(sp, max_len=16, prompt_format=StoredFormat("x"))
    ids = sp.encode("move left 2 cm")
    assert t.decode(np.asarray(ids + [10_000, -1])) == "move left 2 cm"
    with pytest.raises(ValueError, match="Unknown prompt format"):
        tk.CoTTokenizer(sp, prompt_format="nope")


def test_prompt_formats_match_reference_source():
    """N2 third slice: every registered prompt format, the state discretisation and the token-piece checkers against the
    reference modules executed from source (tests/golden/make_reference_prompt_golden.py).  Strings compare byte for byte."""
    import gzip, json, random
    from lap_b200 import prompt_format as pf
    from tests.golden.make_reference_prompt_golden import PIECES, PROMPTS, states

    z = json.load(gzip.open(os.path.join(os.path.dirname(__file__), "golden", "reference_prompts.json.gz"), "rt", encoding="utf-8"))
    assert z["prompts"] == PROMPTS and z["pieces"] == PIECES
    formats = {**{f"train/{k}": v for k, v in pf.PROMPT_FORMAT_REGISTRY.items()},
               **{f"pred/{k}": v for k, v in pf.PREDICTION_PROMPT_FORMAT_REGISTRY.items()},
               "vqa/default_vqa": pf.DEFAULT_VQA_PROMPT_FORMAT}
    assert set(formats) == {r["format"] for r in z["rows"]} == set(z["include_state"])
    st = states()
    for r in z["rows"]:
        got = formats[r["format"]].format_prompt(PROMPTS[r["prompt"]], st[r["state"]], r["state_type"],
                                                  time_horizon_seconds=1.3, frame_description=r["frame"])
        assert got == r["text"], (r, got)
    random.seed(11)
    drops = [pf.LAP_PROMPT_FORMAT.format_prompt("pick", st[2] if k % 3 else None, "eef_pose", state_dropout=0.5) for k in range(12)]
    assert drops == z["drops"] and len(set(drops)) == 2
    disc = [pf.discretize_state(s, bins=b, min_dim=m) for b, m in ((256, 10), (1000, 0), (16, 3)) for s in st[1:]]
    assert disc == z["disc"]
    for fname, fmt in formats.items():
        assert fmt.include_state == z["include_state"][fname]
        for kind in ("critical_token_checker", "direction_token_checker"):
            fn = getattr(fmt, kind)
            assert [bool(fn(p)) if fn is not None else None for p in PIECES] == z["checks"][f"{fname}/{kind}"], (fname, kind)
    # the horizon template's unfilled {frame_description} raises like the reference (prompt.py:59)
    with pytest.raises(KeyError):
        pf.PromptFormat(name="h", include_time_horizon=True).format_prompt("x", time_horizon_seconds=1.0)


def test_tokenizing_transforms_match_reference_source():
    """`TokenizePromptAndReasoning` / `DetokenizeReasoning` / `SafeRepackTransform` against the reference classes executed from
    src/lap/transforms.py (training, inference, VQA and prediction samples; fixture: reference_tokenizer.npz `tf/*`)."""
    sentencepiece = pytest.importorskip("sentencepiece")
    from lap_b200 import tokenizer as tk, transforms as T
    from tests.golden.make_reference_tokenizer_golden import REPACK, TRANSFORM_CASES

    gold = os.path.join(os.path.dirname(__file__), "golden")
    z = np.load(os.path.join(gold, "reference_tokenizer.npz"))
    sp = sentencepiece.SentencePieceProcessor(model_file=os.path.join(gold, "tiny_sp.model"))
    for i, c in enumerate(TRANSFORM_CASES):
        t = tk.CoTTokenizer(sp, max_len=c["max_len"])
        res = T.TokenizePromptAndReasoning(t, **c["kw"])(dict(c["data"]))
        assert "\x00".join(sorted(res)).encode() == bytes(z[f"tf/{i}/keys"])
        assert "\x00".join(sorted(k for k, v in res.items() if v is None)).encode() == bytes(z[f"tf/{i}/none"])
        for k, v in res.items():
            if v is not None and k not in c["data"]:
                ref = z[f"tf/{i}/{k}"]
                assert np.asarray(v).dtype == ref.dtype, k
                np.testing.assert_array_equal(np.asarray(v), ref, err_msg=f"{i}/{k}")
        if i == 0:
            det = T.DetokenizeReasoning(t)({"tokens": res["tokenized_prompt"][None].astype(np.int64), "a": 1})
            assert det["a"] == 1 and det["reasoning"].encode() == bytes(z["tf/detok"])
            assert T.DetokenizeReasoning(t)({"a": 1}) == {"a": 1}
    assert repr(T.SafeRepackTransform(REPACK["structure"])(REPACK["data"])).encode() == bytes(z["tf/repack"])
    with pytest.raises(KeyError, match="Missing source paths"):
        T.SafeRepackTransform(REPACK["structure"], strict=True)(REPACK["data"])
    with pytest.raises(ValueError, match="Prompt is required"):
        T.TokenizePromptAndReasoning(tk.CoTTokenizer(sp))({"is_vqa_sample": False})
    with pytest.raises(ValueError, match="State is required"):
        T.TokenizePromptAndReasoning(tk.CoTTokenizer(sp), discrete_state_input=True)({"prompt": "x"})


def test_language_actions_and_policy_io_match_reference_source():
    """N2 fourth slice: action <-> text (`lap_b200.lang_actions`) and `CoTInputs` / `CoTOutputs` (`lap_b200.policy_io`) against
    the reference modules run from source.  The SAME sweep (`run_all`) is executed on both sides; strings compare byte for
    byte, arrays bit for bit (raw bytes), dict key order included."""
    pytest.importorskip("scipy")
    import gzip, json
    from lap_b200 import lang_actions as LA, policy_io as IO
    from tests.golden import make_reference_langaction_golden as G

    ref = json.load(gzip.open(os.path.join(os.path.dirname(__file__), "golden", "reference_langactions.json.gz"), "rt", encoding="utf-8"))
    got = json.loads(json.dumps(G.run_all(None, IO.CoTInputs, IO.CoTOutputs, LA, LA, LA, str)))
    assert set(got) == set(ref)
    for key in ref:
        assert len(got[key]) == len(ref[key]), key
        if isinstance(ref[key], list):
            for i, (a, b) in enumerate(zip(got[key], ref[key])):
                assert a == b, (key, i, a, b)
        assert got[key] == ref[key], key
    # the fixture is not vacuous: moving and idle samples, both frames, masked and unmasked cameras
    flat = json.dumps(ref["inputs"])
    assert '"end-effector frame"' in flat and '"robot base frame"' in flat and "move " in flat
    assert any(ref["idle_generated"]) and not all(ref["idle_generated"])
    with pytest.raises(ValueError, match="Unknown language action format"):
        LA.get_language_action_format("nope")
    with pytest.raises(NotImplementedError):
        IO.CoTInputs(action_dim=7, enable_diverse_questions=True)


def test_policy_and_ar_policy_host_plumbing_with_a_stub_model():
    """`Policy.infer` / `ARPolicy.infer` (openpi policy.py:68-106, policy_adapter.py:26-50) around a stub model: transforms
    run in order, inputs are batched to 1 (None leaves stay None), the text decoded from the model's tokens is parsed back
    into an action, and the request dict is not modified."""
    sentencepiece = pytest.importorskip("sentencepiece")
    from lap_b200 import lang_actions as LA, policy_io as IO, prompt_format as pf, tokenizer as tk, transforms as T
    from lap_b200.policy import ARPolicy, Policy

    sp = sentencepiece.SentencePieceProcessor(model_file=os.path.join(os.path.dirname(__file__), "golden", "tiny_sp.model"))
    tok = tk.CoTTokenizer(sp, max_len=96)
    answer = "move left 2 cm and move up 1 cm"
    seen = {}

    class Stub:
        def sample_tokens(self, rng, obs, max_decoding_steps=8):
            seen["obs"], seen["rng"] = obs, rng
            ids = sp.encode(answer, add_eos=True)
            return torch.tensor([ids + [0] * (max_decoding_steps - len(ids))], dtype=torch.int32)

        def sample_actions(self, rng, obs, **kw):
            seen["kw"] = kw
            return torch.full((1, 10, 32), 0.5)

    stats = {"actions": T.NormStats(mean=np.zeros(7), std=np.ones(7), q01=-np.ones(7), q99=np.ones(7) * 3)}
    tfs = [IO.CoTInputs(action_dim=32), T.TokenizePromptAndReasoning(tok, discrete_state_input=True), T.PadStates(32)]
    base = Policy(Stub(), transforms=tfs, output_transforms=[T.Unnormalize(stats, "bounds_q99")], metadata={"name": "stub"})
    req = {"observation": {"base_0_rgb": np.full((8, 8, 3), 7, np.uint8), "state": np.linspace(-1, 1, 10)}, "prompt": "stack the cups"}
    keys_before = set(req)
    out = base.infer(req, noise=np.zeros((10, 32)))
    assert out["actions"].shape == (10, 32) and seen["kw"]["noise"].shape == (1, 10, 32)
    np.testing.assert_allclose(out["actions"][:, :7], (0.5 + 1) / 2 * (4 + 1e-6) - 1)   # BOUNDS_Q99 inverse on the 7 real dims
    np.testing.assert_array_equal(out["actions"][:, 7:], 0.5)
    assert set(req) == keys_before and "prompt" in req and base.metadata == {"name": "stub"}

    ar = ARPolicy(Policy(Stub(), transforms=tfs, output_transforms=[T.DetokenizeReasoning(tok), IO.CoTOutputs("verbose_with_rotation")]),
                  sample_kwargs=dict(max_decoding_steps=24))
    out = ar.infer(req)
    obs = seen["obs"]
    assert obs.state.shape == (1, 32) and obs.tokenized_prompt.shape == (1, 96) and obs.tokenized_langact_mask is None
    assert obs.image_masks["left_wrist_0_rgb"].tolist() == [False] and obs.image_masks["base_0_rgb"].tolist() == [True]
    assert "State: " in tok.decode(obs.tokenized_prompt[0])      # the discretised state made it into the prompt
    assert out["reasoning"] == answer
    np.testing.assert_allclose(out["actions"], [0.0, 0.02, 0.01, 0, 0, 0])
    assert ar.metadata == {} and out["policy_timing"]["infer_ms"] >= 0
    mv, g = LA.VERBOSE_WITH_ROTATION_FORMAT.parse_language_to_deltas(answer + ", open gripper")
    assert g == 1.0 and mv[1] == 0.02


def test_checkpoint_round_trip_in_reference_layout(tmp_path):
    """`lap_b200.checkpoint` (reference: src/lap/training/checkpoints.py): save -> restore is bit-exact for params, Adam moments
    and EMA; the `params` item follows the `_split_params` convention (EMA weights when EMA is on) and is a reference-layout
    tree; `keep`, `latest_step`, unfinished saves and mismatching models behave.  Host logic only: the flat buffers live on
    the CPU behind a stand-in with the four `LAP` members the module uses."""
    pytest.importorskip("safetensors")
    import shutil
    from lap_b200 import checkpoint as C
    from lap_b200.train import TrainState

    class CpuModel:
        def __init__(self, cfg, seed):
            self.cfg, self.layout = cfg, P.FlatLayout(cfg)
            self.P = torch.randn(self.layout.total, generator=torch.Generator().manual_seed(seed))
            self.refreshed = 0
        def params_reference(self, flat=None):
            eng = {k: v.detach().float().cpu() for k, v in P.engine_from_flat(self.layout, self.P if flat is None else flat).items()}
            return P.engine_to_reference(self.cfg, eng)
        def refresh_compute_copy(self):
            self.refreshed += 1
        def load_params(self, tree):
            C._into_flat(self, self.P, tree)

    cfg = get_config("debug_tiny").model
    def state(seed, ema=True):
        m = CpuModel(cfg, seed)
        g = torch.Generator().manual_seed(seed + 100)
        n = m.layout.total
        return TrainState(step=seed, model=m, mu=torch.randn(n, generator=g), nu=torch.rand(n, generator=g),
                          ema_params=torch.randn(n, generator=g) if ema else None, ema_decay=0.999 if ema else None)

    a = state(7)
    d = C.save_train_state(tmp_path, a, keep=2)
    assert d == tmp_path / "7" and C.latest_step(tmp_path) == 7
    served = C.load_tree(d / "params.safetensors")
    assert {k: tuple(v.shape) for k, v in served.items()} == {k: tuple(s) for k, s in P.reference_shapes(cfg).items()}
    ema_ref = a.model.params_reference(a.ema_params)
    assert all(torch.equal(served[k], ema_ref[k]) for k in served)          # `params` item = EMA weights (checkpoints.py:529-538)
    b = state(1)
    assert C.restore_train_state(tmp_path, b) == 7 and b.step == 7 and b.model.refreshed == 1
    lay = a.model.layout
    same = lambda x, y: all(torch.equal(lay.view(x, n), lay.view(y, n)) for n in lay.shapes)   # (alignment gaps are not state)
    for x, y in ((a.model.P, b.model.P), (a.mu, b.mu), (a.nu, b.nu), (a.ema_params, b.ema_params)):
        assert same(x, y)
    m = CpuModel(cfg, 3)
    assert C.load_served_params(tmp_path, m) == 7 and same(m.P, a.ema_params)   # what serve_policy.py loads
    # no EMA: `params` holds the raw weights, train_state carries none; restoring drops a stale EMA buffer
    c = state(9, ema=False)
    C.save_train_state(tmp_path, c, keep=2)
    assert not (tmp_path / "9" / "train_state" / "params.safetensors").exists()
    e = state(2)
    C.restore_train_state(tmp_path, e, step=9)
    assert same(e.model.P, c.model.P) and e.ema_params is None and e.ema_decay is None
    # keep=2 prunes the oldest; an unfinished save (no manifest / tmp name) is never picked up
    a.step = 11
    C.save_train_state(tmp_path, a, keep=2)
    assert sorted(p.name for p in tmp_path.iterdir()) == ["11", "9"]
    # keep_period: steps divisible by it survive pruning (CheckpointManagerOptions(max_to_keep=1, keep_period=...))
    for st_ in (20, 25, 30):
        a.step = st_
        C.save_train_state(tmp_path, a, keep=1, keep_period=10)
    assert sorted(int(p.name) for p in tmp_path.iterdir()) == [20, 30]
    for st_ in (20, 30):
        shutil.rmtree(tmp_path / str(st_))
    a.step = 11
    C.save_train_state(tmp_path, a)
    (tmp_path / "12.tmp-1").mkdir(); (tmp_path / "13").mkdir()
    assert C.latest_step(tmp_path) == 11
    with pytest.raises(FileNotFoundError):
        C.restore_train_state(tmp_path, e, step=13)
    with pytest.raises(FileNotFoundError):
        C.restore_train_state(tmp_path / "nowhere", e)
    other = TrainState(step=0, model=CpuModel(get_config("debug_small").model, 0), mu=None, nu=None, ema_params=None, ema_decay=None)
    with pytest.raises(ValueError, match="different model"):
        C.restore_train_state(tmp_path, other)
    bad = dict(served); bad.pop(next(iter(bad)))
    with pytest.raises(ValueError, match="missing"):
        C._into_flat(m, m.P, bad)


def test_language_action_round_trip_properties():
    """Size-independent properties of the action <-> text path (no fixture involved):
    (1) base -> end-effector -> base frame is the identity for the default axis convention (frame_transforms.py:22-129);
    (2) text -> deltas inverts deltas -> text up to the rounding the text format applies (1 cm, 10 degrees, gripper bit);
    (3) a chunk is idle exactly when its summed translation rounds below 1 cm and its rotation below 10 degrees."""
    pytest.importorskip("scipy")
    hyp = pytest.importorskip("hypothesis")
    from hypothesis import given, settings, strategies as st_
    from lap_b200 import lang_actions as LA

    f = lambda lo, hi: st_.floats(lo, hi, allow_nan=False, allow_infinity=False, width=64)

    @settings(max_examples=60, deadline=None)
    @given(st_.lists(f(-0.2, 0.2), min_size=3, max_size=3), st_.lists(f(-0.6, 0.6), min_size=3, max_size=3),
           st_.lists(f(-1, 1), min_size=6, max_size=6), f(0, 1))
    def frames(dp, dr, r6, g):
        r6 = np.asarray(r6)
        a1, a2 = r6[:3], r6[3:]
        if np.linalg.norm(a1) < 0.2 or np.linalg.norm(np.cross(a1, a2)) < 0.2 * np.linalg.norm(a1):
            return  # degenerate 6-D rotation (no unique frame)
        state = np.concatenate([[0.1, 0.2, 0.3], r6, [g]])
        a = np.array([*dp, *dr, g])
        back = LA.transform_actions_from_eef_frame(LA.transform_actions_to_eef_frame(a, state, "droid"), state, "droid")[0]
        np.testing.assert_allclose(back, a, atol=1e-9)

    @settings(max_examples=120, deadline=None)
    @given(st_.lists(st_.lists(f(-0.08, 0.08), min_size=7, max_size=7), min_size=1, max_size=5))
    def text(rows):
        a = np.asarray(rows)
        a[:, 3:6] *= 8.0                       # rotations up to ~0.6 rad per row
        a[:, 6] = np.abs(a[:, 6]) * 12.0       # gripper in [0, ~1)
        fmt = LA.VERBOSE_WITH_ROTATION_FORMAT
        s = LA.summarize_numeric_actions(a, fmt.get_sum_decimal(), fmt.include_rotation)
        mv, grip = fmt.parse_language_to_deltas(s)
        tot = a.sum(0)
        assert np.all(np.abs(mv[:3] * 100 - tot[:3] * 100) <= 0.5 + 1e-9), (s, tot)
        # writer and parser agree on the signs ("tilt back" = +pitch, "tilt left" = +roll, counterclockwise = +yaw); the
        # text carries angles to the nearest 10 degrees, so all three come back within 5
        assert np.all(np.abs(np.degrees(mv[3:6]) - np.degrees(tot[3:6])) <= 5.0 + 1e-9), (s, tot)
        assert grip == (1.0 if a[-1, 6] >= 0.5 else 0.0)
        cm = np.round(np.abs(tot[:3] * 100))
        deg = np.array([LA._nearest(abs(v) * 180 / np.pi, 10) for v in tot[3:6]])
        idle = bool(np.sqrt((cm ** 2).sum()) < 1.0 and np.sqrt((deg ** 2).sum()) < 10.0)
        assert LA.is_idle_language_action(s, "0f", True) == idle, (s, cm, deg)

    frames()
    text()


def test_tokenizer_mask_invariants_property():
    """Invariants of `CoTTokenizer.tokenize` over random prompts, reasoning strings and lengths: the valid tokens are a left
    aligned run padded with pad_id; the lang-action span is a contiguous run inside it that starts right after the prompt
    tokens and (when nothing is truncated) decodes back to the cleaned reasoning text and ends in EOS; number / direction masks
    live inside the span; the loss mask only ever removes span positions."""
    sentencepiece = pytest.importorskip("sentencepiece")
    pytest.importorskip("hypothesis")
    from hypothesis import given, settings, strategies as st_
    from lap_b200 import tokenizer as tk

    sp = sentencepiece.SentencePieceProcessor(model_file=os.path.join(os.path.dirname(__file__), "golden", "tiny_sp.model"))
    words = st_.lists(st_.sampled_from(["move", "left", "right", "up", "down", "3", "12", "cm", "and", "open", "gripper",
                                        "the", "cup", "pick", "rotate", "clockwise", "degrees"]), min_size=1, max_size=12)

    @settings(max_examples=80, deadline=None)
    @given(words, st_.one_of(st_.none(), words), st_.integers(8, 160), st_.sampled_from([0.0, 0.5]), st_.integers(0, 2**31 - 1))
    def check(pw, rw, max_len, drop, seed):
        prompt, reasoning = " ".join(pw), None if rw is None else " ".join(rw)
        t = tk.CoTTokenizer(sp, max_len=max_len, prompt_format="lap", reasoning_mask_prob=drop)
        np.random.seed(seed)
        tokens, attn, span, num, dirn, loss = t.tokenize(prompt, reasoning)
        assert tokens.shape == attn.shape == loss.shape == (max_len,) and tokens.dtype == np.int32
        n = int(attn.sum())
        assert attn[:n].all() and not attn[n:].any() and (tokens[n:] == sp.pad_id()).all()
        assert tokens[0] == sp.bos_id()
        if reasoning is None:
            assert span is None and num is None and dirn is None and loss.all()
            return
        idx = np.flatnonzero(span)
        n_prompt = len(sp.encode(t._prompt_format.format_prompt(prompt), add_bos=True))
        if idx.size:
            assert idx[0] == n_prompt and np.array_equal(idx, np.arange(idx[0], idx[-1] + 1)) and idx[-1] < n
        else:
            assert n_prompt >= max_len                      # the prompt alone filled the window
        assert not (num & ~span).any() and not (dirn & ~span).any()
        assert loss[~span].all()                            # dropout only touches lang-action positions
        if drop == 0.0:
            assert loss.all()
        full = n_prompt + len(sp.encode(reasoning, add_eos=True))
        if full <= max_len:
            assert n == full and tokens[n - 1] == sp.eos_id()
            assert sp.decode(tokens[idx].tolist()) == reasoning
            pieces = [sp.id_to_piece(int(v)) for v in tokens[idx]]
            assert num[idx].tolist() == [tk.is_number(p) for p in pieces]
            assert dirn[idx].tolist() == [tk.is_direction_natural(p) for p in pieces]

    check()


def test_ema_schedule_matches_reference_source():
    """`TrainConfig.get_ema_init` / `get_ema_decay_for_step` against the reference's TrainConfig methods and EmaSchedule classes
    executed from src/lap/training/config.py (tests/golden/make_reference_schedule_golden.py): every schedule kind x decay x
    start step x training length, at steps around every boundary."""
    import dataclasses, json
    from lap_b200.config import EmaScheduleChoice
    z = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_schedules.json")))
    base = get_config("lap_libero")
    n_enabled = 0
    for r in z["rows"]:
        c = dataclasses.replace(base, ema_decay=r["decay"], num_train_steps=r["num_train_steps"],
                                ema_schedule_choice=EmaScheduleChoice(kind=r["kind"], start_step=r["start"]))
        d0, e0 = c.get_ema_init()
        assert (None if d0 is None else float(d0)) == r["init"][0] and bool(e0) == r["init"][1], (r["kind"], r["decay"], r["start"])
        for step, (dr, er) in zip(z["steps"], r["steps"]):
            d, e = c.get_ema_decay_for_step(step)
            assert bool(e) == er, (r, step)
            assert abs(float(d) - dr) <= 2e-7 * max(1.0, abs(dr)), (r["kind"], r["decay"], r["start"], step, d, dr)  # reference: fp32
            n_enabled += er
    assert n_enabled > 200


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reads the reference sources (build container only)")
def test_model_dimensions_match_reference_source():
    """The LAP-3B dimensions this engine is built for are the reference's: `gemma.get_config` (backbones/gemma.py:56-110) and
    `siglip.decode_variant` (OP/models/siglip.py:298-370) executed from source.  (Skipped where /root/reference is absent.)"""
    import ast, dataclasses, types
    from lap_b200.config import get_gemma_config, get_siglip_config
    src = "/root/reference/src/lap/models/backbones/gemma.py"
    tree = ast.parse(open(src).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "get_config")
    fn.returns = None
    fn.args.args[0].annotation = None
    ns = {"Config": lambda **kw: types.SimpleNamespace(**kw), "lora": types.SimpleNamespace(LoRAConfig=lambda **kw: kw)}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), src, "exec"), ns)
    for variant in ("dummy", "gemma_300m", "gemma_2b"):
        r, m = ns["get_config"](variant), get_gemma_config(variant)
        for f in ("width", "depth", "mlp_dim", "num_heads", "num_kv_heads", "head_dim"):
            assert getattr(r, f) == getattr(m, f), (variant, f)
    src = "/root/reference/third_party/openpi/src/openpi/models/siglip.py"
    tree = ast.parse(open(src).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "decode_variant")
    ns = {}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), src, "exec"), ns)
    r, m = ns["decode_variant"]("So400m/14"), get_siglip_config("So400m/14", 2048)
    assert (r["width"], r["depth"], r["mlp_dim"], r["num_heads"], r["patch_size"]) == (m.width, m.depth, m.mlp_dim, m.num_heads, (m.patch_size, m.patch_size))


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reads the reference sources (build container only)")
def test_lap_config_defaults_match_reference_source():
    """Every field `LAPConfig` shares with the reference dataclass (src/lap/models/lap_config.py:22-75) has the same default;
    the reference fields this engine does not carry are exactly the ones for out-of-scope heads."""
    import ast, dataclasses
    tree = ast.parse(open("/root/reference/src/lap/models/lap_config.py").read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "LAPConfig")
    mine = {f.name: f.default for f in dataclasses.fields(LAPConfig)}
    absent = []
    for n in cls.body:
        if isinstance(n, ast.AnnAssign) and n.value is not None:
            try:
                ref_default = ast.literal_eval(n.value)
            except ValueError:
                continue
            if n.target.id in mine:
                assert mine[n.target.id] == ref_default, n.target.id
            else:
                absent.append(n.target.id)
    assert sorted(absent) == ["use_fast", "vqa_loss_weights"], absent
    sentencepiece = pytest.importorskip("sentencepiece")
    sp = sentencepiece.SentencePieceProcessor(model_file=os.path.join(os.path.dirname(__file__), "golden", "tiny_sp.model"))
    t = LAPConfig(max_token_len=33, reasoning_mask_prob=0.25).make_tokenizer(sp)
    assert t._max_len == 33 and t.reasoning_mask_prob == 0.25 and t._prompt_format.name == "lap"


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reads the reference sources (build container only)")
def test_train_configs_match_reference_source():
    """`get_config("lap")` / `get_config("lap_libero")` and the TrainConfig defaults against src/lap/training/config.py, read
    through its AST: every keyword the reference passes (and every default it leaves) that this engine's TrainConfig carries."""
    import ast, dataclasses
    tree = ast.parse(open("/root/reference/src/lap/training/config.py").read())
    lit = ast.literal_eval

    def kwargs(call):
        return {k.arg: k.value for k in call.keywords}

    # class defaults
    tc = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "TrainConfig")
    ref_default = {}
    for n in tc.body:
        if isinstance(n, ast.AnnAssign) and n.value is not None:
            try:
                ref_default[n.target.id] = lit(n.value)
            except ValueError:
                pass
    ema_def = next(n for n in tc.body if isinstance(n, ast.AnnAssign) and n.target.id == "ema_schedule_choice")
    ema_call = next(c for c in ast.walk(ema_def.value) if isinstance(c, ast.Call) and getattr(c.func, "id", "") == "EmaScheduleChoice")
    ref_default["ema_schedule_choice"] = {k: lit(v) for k, v in kwargs(ema_call).items()}
    lr_fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "build_cosine_lr")
    ref_default["lr_schedule"] = {a.arg: lit(d) for a, d in zip(lr_fn.args.kwonlyargs, lr_fn.args.kw_defaults)}
    ref_default["batch_size"] = 32   # inherited: OP/training/config.py:497
    mine_default = TrainConfig()
    for f in ("num_train_steps", "save_interval", "log_interval", "keep_period", "seed", "ema_decay", "batch_size"):
        assert getattr(mine_default, f) == ref_default[f], f
    assert dataclasses.asdict(mine_default.lr_schedule) == ref_default["lr_schedule"]
    assert dataclasses.asdict(mine_default.ema_schedule_choice) == ref_default["ema_schedule_choice"]
    assert mine_default.optimizer.weight_decay == 1e-4

    # the two named configurations of the hot path
    cfgs = next(n for n in tree.body if isinstance(n, ast.Assign) and getattr(n.targets[0], "id", None) == "_CONFIGS").value
    seen = set()
    for call in cfgs.elts:
        if not isinstance(call, ast.Call):
            continue
        kw = kwargs(call)
        name = lit(kw["name"])
        if name not in ("lap", "lap_libero"):
            continue
        seen.add(name)
        mine = get_config(name)
        model_kw = {k: lit(v) for k, v in kwargs(kw["model"]).items()}
        for k, v in model_kw.items():
            assert getattr(mine.model, k) == v, (name, k)
        for k, v in dataclasses.asdict(LAPConfig()).items():   # fields the reference leaves at their defaults
            if k not in model_kw:
                assert getattr(mine.model, k) == v, (name, k)
        for f in ("num_train_steps", "save_interval", "keep_period", "batch_size"):
            want = lit(kw[f]) if f in kw else ref_default[f]
            assert getattr(mine, f) == want, (name, f, getattr(mine, f), want)
        want_lr = {k: lit(v) for k, v in kwargs(kw["lr_schedule"]).items()} if "lr_schedule" in kw else ref_default["lr_schedule"]
        assert dataclasses.asdict(mine.lr_schedule) == want_lr, name
        want_ema = dict(ref_default["ema_schedule_choice"])
        if "ema_schedule_choice" in kw:
            want_ema = {**{"kind": "delayed", "start_step": 10000}, **{k: lit(v) for k, v in kwargs(kw["ema_schedule_choice"]).items()}}
        assert dataclasses.asdict(mine.ema_schedule_choice) == want_ema, (name, want_ema)
    assert seen == {"lap", "lap_libero"}


def test_training_loop_resume_save_and_log_cadence(tmp_path):
    """`run_training` (the loop of scripts/train.py:main, :458-620) with a stand-in runner: checkpoints at the reference's
    cadence with its retention, resume continues from the newest checkpoint's step with its state, logs average the infos."""
    pytest.importorskip("safetensors")
    import dataclasses
    from lap_b200 import checkpoint as C
    from lap_b200.train import TrainState
    from lap_b200.train_loop import run_training

    cfg = get_config("debug_tiny")
    cfg = dataclasses.replace(cfg, num_train_steps=23, save_interval=5, keep_period=10, log_interval=4)

    class CpuModel:
        def __init__(self, mc):
            self.cfg, self.layout = mc, P.FlatLayout(mc)
            self.P = torch.zeros(self.layout.total)
        def params_reference(self, flat=None):
            eng = {k: v.detach().float().cpu() for k, v in P.engine_from_flat(self.layout, self.P if flat is None else flat).items()}
            return P.engine_to_reference(self.cfg, eng)
        def refresh_compute_copy(self):
            pass

    def fresh():
        m = CpuModel(cfg.model)
        n = m.layout.total
        return TrainState(step=0, model=m, mu=torch.zeros(n), nu=torch.zeros(n), ema_params=torch.zeros(n), ema_decay=0.999)

    calls = []
    def runner(rng, state, batch, step):        # "training": params count the steps, loss = the batch value
        assert step == state.step
        state.model.P += 1.0
        state.step = step + 1
        calls.append(step)
        return state, {"loss": torch.tensor(float(batch)), "skipped": None}

    logs = []
    s1 = run_training(cfg, iter(range(100, 112)), checkpoint_dir=tmp_path, state=fresh(), runner=runner, log_fn=lambda s, m: logs.append((s, m)))
    assert calls == list(range(12)) and s1.step == 12                       # the data ran out after 12 batches
    # saves at steps 5 and 10 (step % 5 == 0 and step > 0); retention keeps the newest + multiples of keep_period
    assert sorted(int(p.name) for p in tmp_path.iterdir()) == [10]
    assert [s for s, _ in logs] == [0, 4, 8] and logs[1][1] == {"loss": (101 + 102 + 103 + 104) / 4}
    calls.clear()
    s2 = run_training(cfg, iter(range(1000)), checkpoint_dir=tmp_path, state=fresh(), runner=runner)
    # the checkpoint written DURING step 10 holds the state AFTER that step (state.step = 11), as in the reference
    assert calls[0] == 11 and calls[-1] == 22 and s2.step == 23
    assert float(s2.model.P[0]) == 23.0                                     # 11 restored + 12 new steps
    assert sorted(int(p.name) for p in tmp_path.iterdir()) == [10, 20]      # 15 was pruned, 10 and 20 are keep_period steps
    calls.clear()
    run_training(cfg, iter(range(1000)), checkpoint_dir=tmp_path, resume=False, state=fresh(), runner=runner)
    assert calls[0] == 0


def test_create_trained_policy_assembles_the_reference_chain(tmp_path):
    """`policy_config.create_trained_policy(_ar)` (policy_config_adapter.py:86-160): served weights come from the checkpoint's
    `params` item, norm stats from `assets/<id>/norm_stats.json` (`state_eef_pose` read as `state`), and the transform chain
    runs in the reference's order — CoTInputs, Normalize, tokenizer, pad | detokenize, Unnormalize, CoTOutputs."""
    pytest.importorskip("safetensors")
    sentencepiece = pytest.importorskip("sentencepiece")
    import json
    from lap_b200 import checkpoint as C, policy_config as PC, transforms as T
    from lap_b200.policy_io import CoTInputs, CoTOutputs

    sp = sentencepiece.SentencePieceProcessor(model_file=os.path.join(os.path.dirname(__file__), "golden", "tiny_sp.model"))
    import dataclasses
    tc = get_config("debug_tiny")
    tc = dataclasses.replace(tc, model=dataclasses.replace(tc.model, max_token_len=96))
    # a checkpoint directory as lap_b200.checkpoint writes it, plus the assets the training run stored next to it
    ref_tree = {k: torch.full(s, 0.5) for k, s in P.reference_shapes(tc.model).items()}
    (tmp_path / "7").mkdir()
    C.save_tree(tmp_path / "7" / "params.safetensors", ref_tree)
    (tmp_path / "7" / "meta.json").write_text(json.dumps({"format": C.FORMAT, "step": 7}))
    (tmp_path / "assets" / "libero").mkdir(parents=True)
    stats = {"norm_stats": {"state_eef_pose": {"mean": [0.0] * 8, "std": [1.0] * 8, "q01": [-2.0] * 8, "q99": [2.0] * 8},
                            "actions": {"mean": [0.0] * 7, "std": [1.0] * 7, "q01": [-1.0] * 7, "q99": [3.0] * 7, "num_transitions": 5}}}
    (tmp_path / "assets" / "libero" / "norm_stats.json").write_text(json.dumps(stats))

    seen = {}

    class Stub:
        def load_params(self, tree):
            seen["loaded"] = {k: tuple(v.shape) for k, v in tree.items()}
        def sample_actions(self, rng, obs, **kw):
            seen["obs"] = obs
            return torch.full((1, tc.model.action_horizon, tc.model.action_dim), 0.5)
        def sample_tokens(self, rng, obs, **kw):
            ids = sp.encode("move up 3 cm", add_eos=True)
            return torch.tensor([ids + [0] * 8], dtype=torch.int32)

    pol = PC.create_trained_policy(tc, tmp_path, sp_processor=sp, model=Stub(), default_prompt="stack the cups")
    assert seen["loaded"] == {k: tuple(s) for k, s in P.reference_shapes(tc.model).items()}
    kinds = [type(t).__name__ for t in pol._input_transform.__closure__[0].cell_contents]
    assert kinds == ["InjectDefaultPrompt", "CoTInputs", "Normalize", "InjectDefaultPrompt", "TokenizePromptAndReasoning", "PadStatesAndActions"]
    outs = [type(t).__name__ for t in pol._output_transform.__closure__[0].cell_contents]
    assert outs == ["DetokenizeReasoning", "Unnormalize", "CoTOutputs"]
    req = {"observation": {"base_0_rgb": np.full((8, 8, 3), 9, np.uint8), "state": np.full(8, 1.0)}, "prompt": "stack the cups"}
    out = pol.infer(req)
    obs = seen["obs"]
    # state 1.0 under BOUNDS_Q99 with q01 = -2, q99 = 2 -> 0.5 (then discretised into the prompt and padded to action_dim)
    np.testing.assert_allclose(np.asarray(obs.state)[0, :7], 0.5, atol=1e-6)
    tok = pol._input_transform.__closure__[0].cell_contents[4].tokenizer
    text = tok.decode(np.asarray(obs.tokenized_prompt)[0])
    assert "Task: stack the cups" in text and "State: 191 191" in text      # 0.5 -> bin 191 of 256 on [-1, 1)
    # reference quirk kept: an injected default prompt is a 0-d array, which TextParser.decode_text reads as "" (text_utils.py:8-22)
    pol.infer({"observation": req["observation"]})
    assert "Task: , predict" in tok.decode(np.asarray(seen["obs"].tokenized_prompt)[0])
    # actions 0.5 -> (0.5 + 1) / 2 * (3 - -1 + 1e-6) - 1 = 2.0 on the 7 dims the statistics cover
    np.testing.assert_allclose(out["actions"][:, :7], 2.0, atol=1e-5)
    ar = PC.create_trained_policy_ar(tc, tmp_path, sp_processor=sp, model=Stub(), default_prompt="stack the cups",
                                     data=PC.DataConfig(language_action_format_name="verbose_with_rotation"))
    out = ar.infer(req)
    assert out["reasoning"] == "move up 3 cm"
    np.testing.assert_allclose(out["actions"], [0, 0, 0.03, 0, 0, 0])
    # vla0 strategy: CoTOutputs carries the statistics itself, no separate Unnormalize
    _, o2 = PC.policy_transforms(tc.model, tc.model.make_tokenizer(sp), PC.load_norm_stats(tmp_path / "assets"),
                                 data=PC.DataConfig(language_action_format_name="vla0_chunked", transform_strategy="vla0"))
    assert [type(t).__name__ for t in o2] == ["DetokenizeReasoning", "CoTOutputs"] and o2[1].norm_stats is not None
    with pytest.raises(AssertionError, match="exactly one norm stats directory"):
        PC.load_norm_stats(tmp_path)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="executes the reference's torch implementation (build container only)")
def test_resize_with_pad_matches_reference_torch_implementation():
    """`image_tools.resize_with_pad(antialias=False)` against `resize_with_pad_torch` of the reference
    (third_party/openpi/src/openpi/shared/image_tools.py:55-128) compiled from its source: float32 and uint8 images, up- and
    down-scaling, both aspect directions, batched and single images.  With antialiasing (the JAX path's default) upscaling
    is the same function and downscaling keeps constants, the padding and the value range."""
    import ast
    import torch.nn.functional as F
    from lap_b200.image_tools import resize_with_pad
    src = "/root/reference/third_party/openpi/src/openpi/shared/image_tools.py"
    fn = next(n for n in ast.parse(open(src).read()).body if isinstance(n, ast.FunctionDef) and n.name == "resize_with_pad_torch")
    fn.returns = None
    for a in fn.args.args:
        a.annotation = None
    ns = {"torch": torch, "F": F}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), src, "exec"), ns)
    ref = ns["resize_with_pad_torch"]
    rng = np.random.default_rng(0)
    for shape, (h, w) in [((2, 20, 32, 3), (56, 56)), ((2, 96, 60, 3), (56, 56)), ((1, 300, 400, 3), (224, 224)),
                          ((3, 40, 40, 3), (56, 56)), ((2, 56, 56, 3), (56, 56)), ((17, 33, 3), (24, 40))]:
        f32 = rng.uniform(-1, 1, shape).astype(np.float32)
        u8 = rng.integers(0, 256, shape, dtype=np.uint8)
        for img in (f32, u8):
            want = ref(torch.from_numpy(img), h, w).numpy()
            got = resize_with_pad(img, h, w, antialias=False)
            assert got.shape == (*shape[:-3], h, w, 3) and got.dtype == img.dtype
            want = want.reshape(got.shape)   # (the torch version drops a batch dimension of 1; the JAX version keeps it)
            if img.dtype == np.uint8:
                # torch interpolates uint8 tensors with its own integer kernel (its rounding differs from "interpolate in
                # float, then round" of the JAX path, which this function follows): agreement to one grey level
                assert (np.abs(got.astype(int) - want.astype(int)) <= 1).all()
            else:
                np.testing.assert_allclose(got, want, atol=1e-4)   # torch evaluates the source coordinates in fp32 (error grows with size)
            aa = resize_with_pad(img, h, w, antialias=True)
            ratio = max(shape[-2] / w, shape[-3] / h)
            if ratio <= 1.0:            # upscaling: the antialiasing kernel has unit width
                np.testing.assert_array_equal(aa, got)
            assert aa.shape == got.shape and aa.dtype == img.dtype
    const = np.full((1, 300, 400, 3), 0.25, np.float32)
    out = resize_with_pad(const, 224, 224)
    rows = out[0, :, 0, 0]
    assert np.allclose(rows[28:196], 0.25, atol=1e-6) and (rows[:28] == -1).all() and (rows[196:] == -1).all()   # 168 rows + 28 + 28
    with pytest.raises(ValueError, match="Unsupported image dtype"):
        resize_with_pad(const.astype(np.float64), 224, 224)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reads the reference sources (build container only)")
def test_param_norm_filter_matches_reference_source():
    """`param_norm` is the global norm of the KERNEL parameters: nnx.All(Param, Not(PathRegex(<pattern>)), ndim > 1)
    (scripts/train.py:402-409) with PathRegex = re.fullmatch on the '/'-joined path (OP/shared/nnx_utils.py:47-63).  The pattern
    is read from the reference's train.py; the oracle's predicate and the engine's contiguous kernel range [0, kernel_end)
    select exactly those parameters of the LAP-3B tree."""
    import ast, re
    from oracle import lap_oracle as O
    tree = ast.parse(open("/root/reference/scripts/train.py").read())
    call = next(n for n in ast.walk(tree) if isinstance(n, ast.Call) and getattr(n.func, "attr", "") == "PathRegex")
    pattern = re.compile(ast.literal_eval(call.args[0]))
    cfg = get_config("lap_libero").model
    shapes = P.reference_shapes(cfg)
    want = {k for k, s in shapes.items() if len(s) > 1 and pattern.fullmatch(k) is None}
    assert 0 < len(want) < len(shapes)
    got_oracle = {k for k, s in shapes.items() if O.is_kernel_param(k, torch.empty(0).reshape((0,) * len(s)) if len(s) else torch.empty(()))}
    assert got_oracle == want
    # engine: the kernel tensors are one contiguous prefix of the flat buffer; every reference kernel parameter maps into it and
    # nothing else does (element counts agree)
    lay = P.FlatLayout(cfg)
    n_kernel_ref = sum(int(np.prod(shapes[k])) for k in want)
    n_kernel_eng = sum(int(np.prod(lay.shapes[n])) for n in lay.kernel_names)
    assert n_kernel_eng == n_kernel_ref
    assert lay.kernel_end >= n_kernel_eng      # (alignment gaps are zeros and do not change a norm)


def test_product_package_never_touches_the_oracle_or_the_reference():
    """The rule of the tier: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use
    `oracle/`; nothing under lap_b200/ imports it, reads /root/reference, or falls back to a CPU path."""
    import ast, glob
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for path in glob.glob(os.path.join(root, "lap_b200", "*.py")):
        src = open(path).read()
        tree = ast.parse(src)
        for n in ast.walk(tree):
            if isinstance(n, ast.Import):
                assert not any(a.name.split(".")[0] == "oracle" for a in n.names), path
            if isinstance(n, ast.ImportFrom):
                assert (n.module or "").split(".")[0] != "oracle", path
        assert "/root/reference" not in src, path
    # bench.py: the oracle only inside the two CPU-baseline functions
    tree = ast.parse(open(os.path.join(root, "bench.py")).read())
    users = {f.name for f in ast.walk(tree) if isinstance(f, ast.FunctionDef)
             for n in ast.walk(f) if isinstance(n, ast.ImportFrom) and (n.module or "").split(".")[0] == "oracle"}
    assert users and all("cpu" in u or "reference" in u for u in users), users


def test_train_phases_cover_every_gradient_exactly_once():
    """train.py::_phases (the data-parallel schedule): the gradient ranges all-reduced after the phases tile the flat
    buffer — every tensor exactly once, layer groups in backward order — for world 1 and world > 1."""
    from lap_b200 import params as P
    from lap_b200.config import get_config
    from lap_b200.model import LAP
    from lap_b200.train import TrainingStepRunner

    for name in ("debug_tiny", "lap_libero"):
        tc = get_config(name)
        model = LAP.__new__(LAP)
        model.cfg = model.config = tc.model
        model.layout = P.FlatLayout(tc.model)
        lay = model.layout
        for world, segs in ((1, 3), (2, 3), (8, 2), (2, 50)):
            r = TrainingStepRunner(tc)
            r.world, r.bwd_segments, r.vis_segments = world, segs, segs
            phases = r._phases(model)
            L, Ls = tc.model.gemma.depth, tc.model.siglip.depth
            assert len(phases) == (2 if world == 1 else min(segs, L) + min(segs, Ls))
            assert [p[2] for p in phases][0] is False and all(p[2] for p in phases[1:])
            covered = np.zeros(lay.total, dtype=np.int8)
            for _, ranges, _ in phases:
                for lo, hi in ranges:
                    assert 0 <= lo < hi <= lay.total
                    covered[lo:hi] += 1
            assert covered.max() == 1
            for n, shape in lay.shapes.items():
                o = lay.offsets[n]
                assert covered[o:o + int(np.prod(shape))].min() == 1, n


def test_orbax_plain_layout_round_trip(tmp_path):
    """N1: the Orbax `params` item in its plain-directory layout (one zarr-v2 array per leaf, zstd chunks) written and read
    back bit-exactly — nested tree, `params.` prefix, nnx `value` suffix (model.py:326-332), multi-chunk arrays with ragged
    edge chunks, bf16 leaves, scalars; an OCDBT directory is recognised and refused with the conversion hint."""
    import ml_dtypes
    from lap_b200 import orbax_io

    cfg = get_config("debug_tiny").model
    ref = P.init_reference_params(cfg, 3, reference_zero_init=False)
    tree = {k: v.numpy() for k, v in ref.items()}
    for suffix in (False, True):
        d = tmp_path / f"ckpt{int(suffix)}" / "params"
        orbax_io.write_params(d, tree, value_suffix=suffix, max_chunk_bytes=4096)   # forces multi-chunk leaves
        names = sorted(p.name for p in d.iterdir() if p.is_dir())
        assert all(n.startswith("params.") for n in names) and all(n.endswith(".value") == suffix for n in names)
        assert (d / "_METADATA").exists()
        back = P.from_nested(orbax_io.read_params(d))
        assert set(back) == set(tree)
        assert all(np.array_equal(back[k], tree[k]) and back[k].dtype == tree[k].dtype for k in tree)
    # chunk arithmetic on awkward shapes, uncompressed and compressed, other dtypes
    rng = np.random.default_rng(0)
    for a, chunks, comp in ((rng.standard_normal((7, 5, 3)).astype(np.float32), (3, 2, 3), True),
                            (rng.integers(0, 255, (10,), dtype=np.uint8), (4,), False),
                            (rng.standard_normal((4, 6)).astype(ml_dtypes.bfloat16), (4, 6), True),
                            (np.float32(3.5).reshape(()), None, True)):
        orbax_io.write_zarr_array(tmp_path / "arr", a, chunks=chunks, compress=comp)
        b = orbax_io.read_zarr_array(tmp_path / "arr")
        assert b.dtype == a.dtype and b.shape == a.shape and b.tobytes() == np.asarray(a).tobytes()
        for f in (tmp_path / "arr").iterdir():
            f.unlink()
    assert orbax_io.zstd_decompress(orbax_io.zstd_compress(b"lap" * 1000)) == b"lap" * 1000
    oc = tmp_path / "ocdbt" / "params"
    oc.mkdir(parents=True)
    (oc / "manifest.ocdbt").write_bytes(b"\x0c\xdb\x3a\x2a")
    with pytest.raises(NotImplementedError, match="convert_orbax_checkpoint"):
        orbax_io.read_params(oc)


def test_denoise_loop_shape_dispatch_is_host_logic():
    """lapb200_denoise_supported (csrc/denoise.cu) is pure host arithmetic over the two shared-memory layouts of K10: the
    serving shape of LAP-3B (10 action rows, 692 prefix keys) and its 16-row variant (round-1 layout only: the v2 buffers
    no longer fit) are taken by the persistent kernel, the 50-row BASELINE.json variant and batches > 1 are not."""
    from lap_b200 import ops
    e = get_config("lap_libero").model.expert
    dims = (e.width, e.num_heads, e.head_dim, e.mlp_dim)
    assert dims == (1024, 8, 256, 4096)
    ok = lambda B, A, ad, Pn, steps=10: ops.denoise_supported(B, A, ad, *dims, Pn, ((Pn + A + 63) // 64) * 64, steps)
    assert ok(1, 10, 7, 692) and ok(1, 10, 32, 692) and ok(1, 16, 32, 692) and ok(1, 10, 7, 692, steps=16)
    assert not ok(1, 50, 32, 560)        # BJ shape: 50 action rows -> kernel-per-op path
    assert not ok(2, 10, 7, 692)         # the loop kernel is the batch-1 serving path
    assert not ok(1, 10, 7, 692, steps=17)
    assert not ok(1, 10, 7, 1100)        # key padding beyond 1024
