"""Generate tests/golden/reference_schedules.json: the EMA schedule of the reference's TrainConfig
(src/lap/training/config.py: `EmaStage` :372-387, `EmaSchedule` :390-455, `EmaScheduleChoice` :458-504, `TrainConfig.ema_schedule`
/ `get_ema_init` / `get_ema_decay_for_step` :545-589) executed from source (numpy for jax.numpy) over every schedule kind,
decay and start step.  Run: python tests/golden/make_reference_schedule_golden.py"""
import ast
import dataclasses
import json
import os
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("LAP_REFERENCE", "/root/reference")
CONFIG_PY = os.path.join(REF, "src/lap/training/config.py")

KINDS = ["disabled", "constant", "delayed", "cosine_delayed"]
DECAYS = [None, 0.999, 0.99]
STARTS = [0, 1000, 5000]
STEPS = [0, 1, 999, 1000, 1001, 4999, 5000, 12345, 39999, 40000, 40001, 100000]
NUM_TRAIN_STEPS = [40_000, 40_001, 3000]


def load():
    tree = ast.parse(open(CONFIG_PY).read())
    classes = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name in ("EmaStage", "EmaSchedule", "EmaScheduleChoice")]
    tc = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "TrainConfig")
    methods = [n for n in tc.body if isinstance(n, ast.FunctionDef) and n.name in ("ema_schedule", "get_ema_init", "get_ema_decay_for_step")]
    for m in methods:
        m.decorator_list, m.returns = [], None
        # `import jax.numpy as jnp` inside the method bodies -> the numpy module injected below
        m.body = [s for s in m.body if not isinstance(s, ast.Import)]
        for sub in ast.walk(m):
            if isinstance(sub, (ast.If,)):
                sub.body = [s for s in sub.body if not isinstance(s, ast.Import)] or [ast.Pass()]
    for c in classes:
        for sub in ast.walk(c):
            if isinstance(sub, ast.FunctionDef):
                sub.returns = None
                sub.body = [s for s in sub.body if not isinstance(s, ast.Import)]
    ns = {"dataclasses": dataclasses, "Literal": __import__("typing").Literal, "jnp": np}
    src = ast.fix_missing_locations(ast.Module(body=classes + methods, type_ignores=[]))
    exec(compile(src, CONFIG_PY, "exec"), ns)
    return ns


def main():
    ns = load()
    rows = []
    for kind in KINDS:
        for decay in DECAYS:
            for start in STARTS:
                for nts in NUM_TRAIN_STEPS:
                    self = types.SimpleNamespace(ema_decay=decay, num_train_steps=nts,
                                                 ema_schedule_choice=ns["EmaScheduleChoice"](kind=kind, start_step=start))
                    self.ema_schedule = ns["ema_schedule"](self)       # (a property on the real class)
                    init = ns["get_ema_init"](self)
                    per_step = []
                    for step in STEPS:
                        d, e = ns["get_ema_decay_for_step"](self, np.asarray(step))
                        per_step.append([float(np.asarray(d)), bool(np.asarray(e))])
                    rows.append(dict(kind=kind, decay=decay, start=start, num_train_steps=nts,
                                     init=[None if init[0] is None else float(init[0]), bool(init[1])], steps=per_step))
    json.dump(dict(steps=STEPS, rows=rows), open(os.path.join(HERE, "reference_schedules.json"), "w"))
    print(len(rows), "configurations;", rows[-1]["init"], rows[-1]["steps"][:4])


if __name__ == "__main__":
    main()
