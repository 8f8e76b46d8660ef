"""Generate tests/golden/reference_tokenizer.npz: `PaligemmaTokenizer.tokenize` of the reference
(src/lap/models/tokenizer.py:74-315), executed from its source with the reference's REAL "lap" prompt format
(src/lap/models/prompt_utils/{prompt,state,checkers}.py, loaded as plain modules) on a tiny SentencePiece model trained offline
(tests/golden/tiny_sp.model; the PaliGemma model is a download).  The class bodies are compiled out of the AST because the
module imports openpi / transformers at the top; `_tokenizer.PaligemmaTokenizer` (the openpi base class, unused by
`tokenize`) is replaced by `object`, and `__init__` is bypassed to inject the SentencePiece processor.
The fixture stores, per case, the formatted prompt string and every output of `tokenize`.
Run: python tests/golden/make_reference_tokenizer_golden.py   (needs /root/reference)"""
import ast
import importlib.util
import logging
import os
import sys
import types

import numpy as np
import sentencepiece

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("LAP_REFERENCE", "/root/reference")
PU = os.path.join(REF, "src/lap/models/prompt_utils")
TOK = os.path.join(REF, "src/lap/models/tokenizer.py")


def load_prompt_utils():
    for pkg in ("lap", "lap.models", "lap.models.prompt_utils"):
        if pkg not in sys.modules:
            m = types.ModuleType(pkg)
            m.__path__ = []
            sys.modules[pkg] = m
    mods = {}
    for name in ("checkers", "state", "prompt"):
        full = f"lap.models.prompt_utils.{name}"
        spec = importlib.util.spec_from_file_location(full, os.path.join(PU, name + ".py"))
        m = importlib.util.module_from_spec(spec)
        sys.modules[full] = m
        spec.loader.exec_module(m)
        setattr(sys.modules["lap.models.prompt_utils"], name, m)
        mods[name] = m
    return mods


def load_reference_tokenizer(mods):
    tree = ast.parse(open(TOK).read())
    keep = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name in ("BaseTokenizer", "PaligemmaTokenizer")]
    for cls in keep:
        for sub in ast.walk(cls):
            if isinstance(sub, ast.FunctionDef):
                sub.returns = None
                for a in sub.args.args + sub.args.kwonlyargs:
                    a.annotation = None
    ns = {"np": np, "logging": logging, "ABC": object, "abstractmethod": lambda f: f,
          "_tokenizer": types.SimpleNamespace(PaligemmaTokenizer=object), "is_number": mods["checkers"].is_number,
          "PromptFormat": mods["prompt"].PromptFormat}
    # BaseTokenizer(ABC) and PaligemmaTokenizer(object, BaseTokenizer): drop the duplicate `object` base
    for cls in keep:
        cls.bases = [b for b in cls.bases if not (isinstance(b, ast.Attribute) and b.attr == "PaligemmaTokenizer")]
        if cls.name == "BaseTokenizer":
            cls.bases = []
    exec(compile(ast.fix_missing_locations(ast.Module(body=keep, type_ignores=[])), TOK, "exec"), ns)
    return ns["PaligemmaTokenizer"]


CASES = [
    dict(prompt="pick up the red block and place it in the bowl", reasoning="move forward 3 cm and move left 2 cm", max_len=96),
    dict(prompt="open_the drawer\nand take the marker out.", reasoning="move back 12 cm, rotate clockwise 15 degrees\nopen gripper ", max_len=112),
    dict(prompt="stack the cups", reasoning=None, max_len=64),
    dict(prompt="put the spoon on the towel", reasoning="move down 7 cm and rotate counterclockwise 30 degrees and close gripper", max_len=62),  # truncated inside the reasoning
    dict(prompt="put the spoon on the towel", reasoning="move down 7 cm", max_len=40),  # truncated inside the prompt
    dict(prompt="pick up the red block", reasoning="move up 1 cm and move right 4 cm and rotate clockwise 20 degrees", max_len=96, reasoning_mask_prob=0.5, seed=3),
    dict(prompt="stack the cups", reasoning="move right 5 cm", max_len=96, state=np.array([0.1, -0.5, 0.9, 0.0, 0.3, -1.0, 1.0]), state_type="eef_pose"),
]


TRANSFORM_CASES = [
    dict(data=dict(prompt="pick up the red block", language_actions="move forward 3 cm and move left 2 cm",
                   dataset_name="libero 10", is_vqa_sample=False, is_prediction_sample=False, extra=np.arange(3)),
         kw=dict(verbose_mode=True, dataset_name_pad_len=12), max_len=96),
    dict(data=dict(prompt=np.asarray("stack the cups"), is_vqa_sample=False, is_prediction_sample=False,
                   frame_description="camera frame", state=np.array([0.1, -0.5, 0.9, 0.0, 0.3, -1.0, 1.0]),
                   time_horizon_seconds=2.0),
         kw=dict(discrete_state_input=True, dataset_name_pad_len=5), max_len=80),      # inference: no language actions
    dict(data=dict(prompt="what is on the table", language_actions="a red block", dataset_name="a much longer name than pad",
                   is_vqa_sample=True, is_prediction_sample=False),
         kw=dict(verbose_mode=True, dataset_name_pad_len=4), max_len=64),              # VQA: no number/direction masks
    dict(data=dict(prompt="which way does the arm move", language_actions="move right +3", dataset_name="p",
                   is_vqa_sample=False, is_prediction_sample=True, state=np.array([0.2, 0.4])),
         kw=dict(verbose_mode=True, discrete_state_input=True, dataset_name_pad_len=3), max_len=96),
]
REPACK = dict(structure={"images": {"base": ["observation/image", "observation/img"], "wrist": "observation/wrist"},
                         "state": "observation/state", "gone": "observation/none"},
              data={"observation": {"img": 1, "wrist": 2, "state": 3, "unused": 4}})


def reference_transforms(mods, Ref, sp, fmt):
    """TokenizePromptAndReasoning / DetokenizeReasoning / SafeRepackTransform of src/lap/transforms.py executed from source."""
    import dataclasses
    sys.path.insert(0, HERE)
    import make_reference_transforms_golden as G
    ns = G.load_reference()
    ns.update(PaligemmaTokenizer=Ref)
    exec(compile(G.defs(G.LAP_T, ["TokenizePromptAndReasoning", "DetokenizeReasoning", "SafeRepackTransform"]), G.LAP_T, "exec"), ns)
    out = {}
    for i, c in enumerate(TRANSFORM_CASES):
        t = Ref.__new__(Ref)
        t._tokenizer, t._max_len, t.reasoning_mask_prob = sp, c["max_len"], 0.0
        t._prompt_format, t._prediction_format = fmt, mods["prompt"].PREDICTION_PROMPT_FORMAT_REGISTRY["default"]
        t._vqa_format = mods["prompt"].DEFAULT_VQA_PROMPT_FORMAT
        res = ns["TokenizePromptAndReasoning"](t, **c["kw"])(dict(c["data"]))
        out[f"tf/{i}/keys"] = np.frombuffer("\x00".join(sorted(res)).encode(), dtype=np.uint8)
        for k, v in res.items():
            if v is not None and k not in c["data"]:
                out[f"tf/{i}/{k}"] = np.asarray(v)
        out[f"tf/{i}/none"] = np.frombuffer("\x00".join(sorted(k for k, v in res.items() if v is None)).encode(), dtype=np.uint8)
        if i == 0:
            det = ns["DetokenizeReasoning"](t)({"tokens": res["tokenized_prompt"][None].astype(np.int64), "a": 1})
            out["tf/detok"] = np.frombuffer(det["reasoning"].encode(), dtype=np.uint8)
    rp = ns["SafeRepackTransform"](REPACK["structure"])(REPACK["data"])
    out["tf/repack"] = np.frombuffer(repr(rp).encode(), dtype=np.uint8)
    return out


def main():
    mods = load_prompt_utils()
    fmt = mods["prompt"].PROMPT_FORMAT_REGISTRY["lap"]
    Ref = load_reference_tokenizer(mods)
    sp = sentencepiece.SentencePieceProcessor(model_file=os.path.join(HERE, "tiny_sp.model"))
    out = {"n_cases": np.int64(len(CASES))}
    for i, c in enumerate(CASES):
        t = Ref.__new__(Ref)
        t._tokenizer, t._max_len = sp, c["max_len"]
        t.reasoning_mask_prob = c.get("reasoning_mask_prob", 0.0)
        t._prompt_format = t._prediction_format = t._vqa_format = fmt
        kw = dict(state=c.get("state"), state_type=c.get("state_type"))
        formatted = fmt.format_prompt(c["prompt"], kw["state"], kw["state_type"], time_horizon_seconds=None,
                                      frame_description="robot base frame", state_dropout=0.0)
        if "seed" in c:
            np.random.seed(c["seed"])
        res = t.tokenize(c["prompt"], c["reasoning"], **kw)
        out[f"{i}/formatted"] = np.frombuffer(formatted.encode(), dtype=np.uint8)
        for name, v in zip(("tokens", "attn", "reasoning", "number", "direction", "loss"), res):
            if v is not None:
                out[f"{i}/{name}"] = np.asarray(v)
        pieces = "\x00".join(p for p in ("right", "-", "+3", "back", "7", "cm") if fmt.direction_token_checker(p))
        out[f"{i}/direction_pieces"] = np.frombuffer(pieces.encode(), dtype=np.uint8)
    out.update(reference_transforms(mods, Ref, sp, fmt))
    np.savez_compressed(os.path.join(HERE, "reference_tokenizer.npz"), **out)
    print({k: (v.shape if hasattr(v, "shape") else v) for k, v in list(out.items())[:9]})


if __name__ == "__main__":
    main()
