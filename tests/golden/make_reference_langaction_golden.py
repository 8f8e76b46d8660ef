"""Generate tests/golden/reference_langactions.json.gz by running the reference's own modules
(src/lap/policies/lang_action_formats.py, question_types.py, transforms/{action_text,frame_transforms,action_processor,
image_utils,image_handler,text_utils,sample_handlers,input_transforms,output_transforms}.py) loaded as real modules from
/root/reference.  Only their heavyweight imports are stubbed: `lap.models.model_adapter` (jax/flax; provides IMAGE_KEYS and
ExtendedModelType, restated below from model_adapter.py:18-34), `lap.datasets.utils.helpers.ActionEncoding` (tensorflow),
`openpi.transforms` (`DataTransformFn` protocol + `pad_to_dim`, compiled from its source) and `openpi.models.model.ModelType`.
Run: python tests/golden/make_reference_langaction_golden.py"""
import ast
import enum
import gzip
import importlib.util
import json
import os
import random
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("LAP_REFERENCE", "/root/reference")
SRC = os.path.join(REF, "src")


def _pkg(name):
    if name not in sys.modules:
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
    return sys.modules[name]


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    parent, _, leaf = name.rpartition(".")
    setattr(_pkg(parent), leaf, m)
    return m


def load_reference():
    for p in ("lap", "lap.models", "lap.datasets", "lap.datasets.utils", "lap.policies", "lap.policies.transforms", "openpi",
              "openpi.models"):
        _pkg(p)

    class ModelType(str, enum.Enum):
        PI0 = "pi0"
        PI0_FAST = "pi0_fast"
        PI05 = "pi05"

    class ExtendedModelType(str, enum.Enum):
        PI0 = "pi0"
        PI0_FAST = "pi0_fast"
        PI05 = "pi05"
        LAP = "lap"
        LAP_FAST = "lap_fast"

    class ActionEncoding(enum.IntEnum):
        EEF_POS = 1
        JOINT_POS = 2
        JOINT_POS_BIMANUAL = 3
        EEF_R6 = 4
        ABS_EEF_POS = 5

    sys.modules["openpi.models.model"] = types.SimpleNamespace(ModelType=ModelType)
    sys.modules["lap.models.model_adapter"] = types.SimpleNamespace(IMAGE_KEYS=("base_0_rgb", "left_wrist_0_rgb"),
                                                                    ExtendedModelType=ExtendedModelType)
    sys.modules["lap.datasets.utils.helpers"] = types.SimpleNamespace(ActionEncoding=ActionEncoding)
    op_t = os.path.join(REF, "third_party/openpi/src/openpi/transforms.py")
    tree = ast.parse(open(op_t).read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "pad_to_dim"]
    ns = {"np": np}
    exec(compile(ast.Module(body=fn, type_ignores=[]), op_t, "exec"), ns)
    sys.modules["openpi.transforms"] = types.SimpleNamespace(DataTransformFn=object, pad_to_dim=ns["pad_to_dim"])
    sys.modules["openpi"].transforms = sys.modules["openpi.transforms"]
    P = os.path.join(SRC, "lap/policies")
    T = os.path.join(P, "transforms")
    mods = {}
    for name in ("frame_transforms", "action_text", "image_utils", "text_utils"):
        mods[name] = _load(f"lap.policies.transforms.{name}", os.path.join(T, name + ".py"))
    mods["lang_action_formats"] = _load("lap.policies.lang_action_formats", os.path.join(P, "lang_action_formats.py"))
    mods["question_types"] = _load("lap.policies.question_types", os.path.join(P, "question_types.py"))
    for name in ("action_processor", "image_handler", "sample_handlers", "input_transforms", "output_transforms"):
        mods[name] = _load(f"lap.policies.transforms.{name}", os.path.join(T, name + ".py"))
    return mods, ExtendedModelType


# ---------------------------------------------------------------- cases (shared with the test)
def action_chunks():
    rng = np.random.default_rng(21)
    out = []
    for k in range(40):
        T = int(rng.integers(1, 6))
        a = np.zeros((T, 7))
        a[:, :3] = rng.normal(0, [0.03, 0.0005, 0.01][k % 3], (T, 3))
        a[:, 3:6] = rng.normal(0, [0.2, 0.01, 0.05][k % 3], (T, 3))
        a[:, 6] = rng.random(T)
        if k % 7 == 0:
            a[:, rng.integers(0, 6)] = 0.0
        out.append(a if k % 5 else a[0])
    out.append(np.zeros((2, 7)))
    out.append(np.array([0.005, -0.005, 0.0151, 0.0, 0.0, 0.0, 0.5]))
    out.append(np.zeros((2, 5)))
    return out


def states():
    rng = np.random.default_rng(22)
    out = []
    for _ in range(6):
        s = np.zeros(10)
        s[:3] = rng.normal(0, 0.3, 3)
        s[3:9] = rng.normal(0, 1.0, 6)
        s[9] = rng.random()
        out.append(s)
    return out


DATASETS = ["libero_10", "jaco_play", "berkeley_autolab_ur5", "fmb", "utaustin_mutex", "viola", "droid",
            "furniture_bench_dataset_converted_externally_to_rlds"]
SUM_DECIMALS = ["0f", "1f", "2f", "no_number", "nearest_10", "compact"]
TEXTS = ["move forward 3 cm, move up 1 cm, move left 2 cm, tilt left 10 degrees, rotate clockwise 20 degrees, open gripper",
         "move back 0 cm, close gripper", "move forward slightly and move down a lot, tilt forward moderately", "",
         "Move Right 12.5 cm, tilt up 30 degrees, tilt back 5 degrees, set gripper to 0.3", "move left", "open gripper",
         "<+03 -01 +00 +05 +00 -10 1>", "<+00 +00 +00 0>", "<+00 -00 +01 +00 +00 +05 0>", "move backward 0.4 cm, tilt down 9 degrees",
         "move up 8 cm, rotate counterclockwise 29.5 degrees", "move down 3.0 cm, tilt right 10 degrees",
         "523 127 890 512 512 512 500", "1 2 x", "move forward 2 cm; move forward 1.5 cm close gripper"]


def image_set():
    rng = np.random.default_rng(23)
    return dict(u8=rng.integers(0, 256, (8, 8, 3), dtype=np.uint8), f32=rng.random((8, 8, 3)).astype(np.float32),
                chw=rng.random((3, 8, 8)), zeros=np.zeros((8, 8, 3), np.uint8), tchw=rng.integers(0, 256, (2, 3, 8, 8), dtype=np.uint8))


def input_cases():
    im = image_set()
    st = states()
    ch2 = action_chunks()
    ch = [np.atleast_2d(a)[0] for a in ch2]     # the end-effector-frame formats take one summed delta per sample
    base = dict(observation={"base_0_rgb": im["u8"], "left_wrist_0_rgb": im["f32"], "state": st[0][:8]}, prompt=b"pick up the block")
    cases = [
        dict(kw={}, data=dict(base)),                                                              # serving request
        dict(kw={}, data=dict(base, frame_description=b"camera frame", dataset_name=b"r1_lite_x", prompt="a@b@do it")),
        dict(kw={}, data=dict(base, language_actions=ch[3], raw_state=st[1], dataset_name="libero_10", actions=np.ones((4, 7)),
                              has_wrist_image=True, time_horizon_seconds=1.5)),
        dict(kw=dict(language_action_format="verbose_with_rotation"), data=dict(base, language_actions=ch2[2], raw_state=st[2],
                                                                               dataset_name="droid", rotation_applied=True)),
        dict(kw=dict(use_rough_scale=True), data=dict(base, language_actions=ch[9], raw_state=st[3], dataset_name="fmb")),
        dict(kw={}, data=dict(base, language_actions=ch[40], raw_state=st[3], dataset_name="viola")),   # idle -> sample_mask False
        dict(kw={}, data=dict(base, language_actions=np.concatenate([ch[1], ch[1]], -1), raw_state=st[1], is_bimanual=True)),
        dict(kw={}, data=dict(base, language_actions=ch[6], raw_state=st[1], is_navigation=True)),
        dict(kw=dict(enable_langact_training=False), data=dict(base, language_actions=ch[6], raw_state=st[1])),
        dict(kw=dict(wrist_image_dropout_prob=0.5, random_mask_prob=0.5, random_base_prob=0.5), seed=4,
             data=dict(base, language_actions=ch[7], raw_state=st[4], has_wrist_image=True, dataset_name="jaco_play")),
        dict(kw=dict(wrist_image_dropout_prob=0.5, random_mask_prob=0.5, random_base_prob=0.5), seed=5,
             data=dict(base, language_actions=ch[8], raw_state=st[4], has_wrist_image=True, dataset_name="x")),
        dict(kw={}, data=dict(observation={"base_0_rgb": "", "left_wrist_0_rgb": im["chw"], "state": st[0]}, prompt="p")),
        dict(kw={}, data=dict(observation={"base_0_rgb": im["u8"], "state": st[0]}, prompt="p")),   # no wrist camera
        dict(kw={}, data=dict(base, is_vqa_sample=True, caption=b"a red block", vqa_dataset_id=3)),
        dict(kw={}, data=dict(base, is_vqa_sample=True)),
        dict(kw={}, data=dict(base, is_prediction_sample=True, language_actions=ch[9], raw_state=st[5])),
        dict(kw={}, data=dict(observation={"left_wrist_0_rgb": im["u8"], "state": st[0]}, prompt="p", is_prediction_sample=True,
                              pred_use_primary=True)),
        dict(kw=dict(model_type="lap_fast"), data=dict(observation={"base_0_rgb": im["zeros"], "state": st[0]}, prompt="p")),
        dict(kw=dict(language_action_format="vla0_chunked", transform_strategy="vla0"),
             data=dict(base, actions=np.linspace(-1.2, 1.2, 70).reshape(10, 7))),
        dict(kw=dict(language_action_format="vla0_chunked", transform_strategy="vla0"), data=dict(base)),
    ]
    return cases


def output_cases():
    st = states()
    stats = types.SimpleNamespace(q01=np.linspace(-1, 0, 7), q99=np.linspace(0.5, 2, 7), min=np.linspace(-2, -1, 5),
                                  max=np.linspace(1, 3, 5), mean=np.linspace(0, 1, 7), std=np.linspace(0.1, 0.7, 7))
    c = [dict(kw={}, data=dict(actions=[[1.0, 2.0]]))]
    for t in TEXTS[:7] + TEXTS[10:13]:
        c.append(dict(kw=dict(language_action_format="verbose_with_rotation"), data=dict(actions=None, reasoning=t)))
        c.append(dict(kw=dict(language_action_format="verbose_eef_with_rotation"), data=dict(actions=None, reasoning=t, raw_state=st[2])))
    c.append(dict(kw=dict(language_action_format="verbose_eef_with_rotation"), data=dict(actions=None, reasoning=TEXTS[0])))
    c.append(dict(kw=dict(language_action_format="verbose_eef_with_rotation"),
                  data=dict(actions=None, reasoning=TEXTS[0], raw_state=np.array([[0.1, 0.2, 0.3, 0.4, -0.5, 0.6, 1.0]]))))
    for nt in ("bounds_q99", "bounds", "normal", "other"):
        c.append(dict(kw=dict(language_action_format="vla0_chunked", transform_strategy="vla0", normalization_type=nt),
                      stats=True, data=dict(actions=None, reasoning=" ".join(str(v) for v in range(0, 1000, 13)))))
    c.append(dict(kw=dict(language_action_format="vla0_chunked", transform_strategy="vla0"), data=dict(actions=None, reasoning=TEXTS[13])))
    c.append(dict(kw=dict(language_action_format="vla0_chunked", transform_strategy="vla0"), data=dict(actions=None, reasoning=TEXTS[14])))
    c.append(dict(kw=dict(language_action_format="verbose_with_rotation", transform_strategy="vla0"), data=dict(actions=None, reasoning=TEXTS[0])))
    return c, stats


def enc(v):
    """JSON-able, exact: arrays as {dtype, shape, hex of the raw bytes}."""
    if isinstance(v, np.ndarray):
        return {"__nd__": str(v.dtype), "shape": list(v.shape), "hex": np.ascontiguousarray(v).tobytes().hex()}
    if isinstance(v, (np.bool_, bool)):
        return bool(v)
    if isinstance(v, (np.integer,)):
        return int(v)
    if isinstance(v, (np.floating, float)):
        return {"__f__": float(v).hex()}
    if isinstance(v, dict):
        return {"__d__": [[k, enc(x)] for k, x in v.items()]}
    if isinstance(v, (list, tuple)):
        return [enc(x) for x in v]
    if isinstance(v, bytes):
        return {"__b__": v.decode()}
    return v


def run_all(S, I, O, F, AT, FT, model_type):
    """The sweep, written once against a namespace of callables: the generator passes the reference's, the test this repo's."""
    out = {}
    chunks, sts = action_chunks(), states()
    out["summaries"] = [[AT.summarize_numeric_actions(a, sd, rot) for a in chunks] for sd in SUM_DECIMALS for rot in (False, True)]
    out["nav"] = [AT.summarize_numeric_actions(a, "nearest_10", include_rotation=True, rotation_precision=10) for a in chunks]
    bi = [np.concatenate([np.atleast_2d(a), np.atleast_2d(b)[: np.atleast_2d(a).shape[0]]], -1)
          for a, b in zip(chunks[:12], chunks[12:24]) if np.atleast_2d(b).shape[0] >= np.atleast_2d(a).shape[0]]
    out["bimanual"] = [[AT.summarize_bimanual_numeric_actions(a, sd, rot) for a in bi + [chunks[0]]]
                       for sd in ("0f", "compact") for rot in (False, True)]
    out["idle"] = [[bool(AT.is_idle_language_action(t, sd, rot)) for t in TEXTS + [None]] for sd in SUM_DECIMALS for rot in (False, True)]
    flat = [s for grp in out["summaries"] for s in grp if s is not None]
    out["idle_generated"] = [bool(AT.is_idle_language_action(t, "0f", True)) for t in flat]
    out["scale"] = [AT.describe_language_action_scale(t) for t in TEXTS + flat[:80] + [None, 3, "  "]]
    out["to_eef"] = [enc(FT.transform_actions_to_eef_frame(np.atleast_2d(chunks[i])[0], sts[i % 6], d, bool(i % 2)))
                     for i in range(16) for d in DATASETS]
    out["from_eef"] = [enc(FT.transform_actions_from_eef_frame(chunks[i], sts[i % 6], d)) for i in range(16) for d in DATASETS]
    out["from_eef_euler"] = enc(FT.transform_actions_from_eef_frame(chunks[1][:, :3], sts[0][:7][None], ""))
    out["rot6d"] = enc(FT.rot6d_to_rotmat(np.stack(sts)[:, 3:9]))
    out["parse"] = []
    for name in ("verbose_with_rotation", "verbose_eef_with_rotation", "vla0_chunked"):
        fmt = F.get_language_action_format(name)
        for t in TEXTS + flat[:40]:
            for st in (None, sts[3]):
                if not t and name != "vla0_chunked":
                    pass
                out["parse"].append(enc(list(fmt.parse_language_to_deltas(t, initial_state=st))))
    compact = F.LanguageActionFormat(name="c", style="compact", include_rotation=True)
    out["parse_compact"] = [enc(list(compact.parse_language_to_deltas(t))) for t in TEXTS]
    out["sum_decimal"] = [f.get_sum_decimal() for f in (compact, F.get_language_action_format("verbose_with_rotation"),
                                                        F.get_language_action_format("vla0_chunked"),
                                                        F.LanguageActionFormat(name="d", decimal_places=2))]
    v0 = F.get_language_action_format("vla0_chunked")
    out["vla0"] = [v0.summarize_actions(np.linspace(-1.3, 1.3, 21).reshape(3, 7)), v0.summarize_actions(np.linspace(-1, 1, 7)),
                   enc(v0.parse_to_full_actions("1 2 3")), enc(v0.parse_to_full_actions(["500 600", "700"])),
                   enc(v0.parse_to_full_actions("a b")), enc(v0.parse_to_full_actions(""))]
    out["inputs"] = []
    for c in input_cases():
        kw = dict(c["kw"])
        if "model_type" in kw:
            kw["model_type"] = model_type(kw["model_type"])
        if "seed" in c:
            np.random.seed(c["seed"])
            random.seed(c["seed"])
        out["inputs"].append(enc(I(action_dim=32, **kw)(dict(c["data"]))))
    cases, stats = output_cases()
    out["outputs"] = [enc(O(**c["kw"], **({"norm_stats": {"actions": stats}} if c.get("stats") else {}))(dict(c["data"]))) for c in cases]
    return out


def main():
    mods, EMT = load_reference()
    out = run_all(None, mods["input_transforms"].CoTInputs, mods["output_transforms"].CoTOutputs, mods["lang_action_formats"],
                  mods["action_text"], mods["frame_transforms"], EMT)
    with gzip.open(os.path.join(HERE, "reference_langactions.json.gz"), "wt", encoding="utf-8") as f:
        json.dump(out, f)
    print({k: len(v) for k, v in out.items()})
    print(out["summaries"][1][:3], out["inputs"][2]["__d__"][-3:])


if __name__ == "__main__":
    main()
