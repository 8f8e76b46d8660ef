"""Generate tests/golden/reference_prompts.json.gz: the strings `PromptFormat.format_prompt` of the reference
(src/lap/models/prompt_utils/prompt.py, state.py — plain Python, loaded as modules from /root/reference) returns for every
registered format over a sweep of prompts and states, plus `discretize_state` on edge values and the checker verdicts on a
list of token pieces.  Run: python tests/golden/make_reference_prompt_golden.py"""
import gzip
import json
import os
import random

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

PROMPTS = ["pick up the red block", " open_the drawer\nand take the marker out. ", "Stack the cups...", ""]
PIECES = ["right", "▁Left", "-", "+3", "back", "7", "cm", "▁counterclockwise", "upward", "x", "▁12", ""]
FRAMES = ["robot base frame", "camera frame"]


def states():
    rng = np.random.default_rng(5)
    out = [None, np.zeros(32), rng.uniform(-1, 1, 7), np.concatenate([rng.uniform(-1, 1, 12), np.zeros(20)]),
           np.array([-1.0, 1.0, -1.5, 1.5, 0.0, 1e-9, 0.9921875, 0.99218, -0.0078125, 0.5, 0.25, 0.0]),
           rng.uniform(-1, 1, (2, 8)), np.float32(rng.uniform(-1, 1, 16)), np.zeros((0,))]
    return out


def main():
    import sys
    sys.path.insert(0, HERE)
    from make_reference_tokenizer_golden import load_prompt_utils
    mods = load_prompt_utils()
    P = mods["prompt"]
    formats = {**{f"train/{k}": v for k, v in P.PROMPT_FORMAT_REGISTRY.items()},
               **{f"pred/{k}": v for k, v in P.PREDICTION_PROMPT_FORMAT_REGISTRY.items()},
               "vqa/default_vqa": P.DEFAULT_VQA_PROMPT_FORMAT}
    rows = []
    for fname, fmt in formats.items():
        for pi, prompt in enumerate(PROMPTS):
            for si, st in enumerate(states()):
                for state_type in (None, "eef_pose", "joint_pos", "none", "custom"):
                    frame = FRAMES[(pi + si) % 2]
                    text = fmt.format_prompt(prompt, st, state_type, time_horizon_seconds=1.3, frame_description=frame)
                    rows.append(dict(format=fname, prompt=pi, state=si, state_type=state_type, frame=frame, text=text))
    # state dropout draws from `random` only when a state would be added
    drops = []
    fmt = P.PROMPT_FORMAT_REGISTRY["lap"]
    random.seed(11)
    for k in range(12):
        drops.append(fmt.format_prompt("pick", states()[2] if k % 3 else None, "eef_pose", state_dropout=0.5))
    disc = [mods["state"].StateDiscretizationConfig(bins=b, min_dim=m).discretize_state(s)
            for b, m in ((256, 10), (1000, 0), (16, 3)) for s in states()[1:]]
    checks = {f"{fname}/{kind}": [bool(getattr(fmt, kind)(p)) if getattr(fmt, kind) is not None else None for p in PIECES]
              for fname, fmt in formats.items() for kind in ("critical_token_checker", "direction_token_checker")}
    include_state = {fname: bool(fmt.include_state) for fname, fmt in formats.items()}
    with gzip.open(os.path.join(HERE, "reference_prompts.json.gz"), "wt", encoding="utf-8") as f:
        json.dump(dict(prompts=PROMPTS, pieces=PIECES, rows=rows, drops=drops, disc=disc, checks=checks,
                       include_state=include_state), f)
    print(len(rows), "rows;", rows[5]["text"])


if __name__ == "__main__":
    main()
