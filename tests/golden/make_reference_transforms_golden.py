"""Generate tests/golden/reference_transforms.npz by executing the reference's Normalize / Unnormalize / PadStates /
pad_to_dim (src/lap/transforms.py:150-289,554-562) and apply_tree / flatten_dict / _assert_quantile_stats
(third_party/openpi/src/openpi/transforms.py) from their source files.  The modules import jax/flax at the top, so the class
and function definitions are compiled out of the AST; the only substitutions are `traverse_util.flatten_dict/unflatten_dict`
(flax, third-party: '/'-joined keys of a nested dict) and a plain dataclass for `NormStats` (+ min/max, which the
reference's BOUNDS branch reads).  Run: python tests/golden/make_reference_transforms_golden.py  (needs /root/reference)."""
import ast
import dataclasses
import enum
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("LAP_REFERENCE", "/root/reference")
LAP_T = os.path.join(REF, "src/lap/transforms.py")
OP_T = os.path.join(REF, "third_party/openpi/src/openpi/transforms.py")
HELPERS = os.path.join(REF, "src/lap/datasets/utils/helpers.py")


def defs(path, names):
    tree = ast.parse(open(path).read())
    got = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in names]
    assert {g.name for g in got} == set(names), (path, set(names) - {g.name for g in got})
    for n in got:
        for sub in ast.walk(n):
            if isinstance(sub, ast.FunctionDef):
                sub.returns = None
                for a in sub.args.args + sub.args.kwonlyargs:
                    a.annotation = None
            if isinstance(sub, ast.AnnAssign) and not isinstance(n, ast.FunctionDef):
                sub.annotation = ast.Name(id="object", ctx=ast.Load())
    return ast.fix_missing_locations(ast.Module(body=got, type_ignores=[]))


@dataclasses.dataclass
class NormStats:
    mean: np.ndarray
    std: np.ndarray
    q01: np.ndarray = None
    q99: np.ndarray = None
    min: np.ndarray = None
    max: np.ndarray = None


def _flatten(tree, sep="/"):
    out = {}

    def rec(prefix, node):
        if isinstance(node, dict) and node:
            for k, v in node.items():
                rec(f"{prefix}{sep}{k}" if prefix else str(k), v)
        else:
            out[prefix] = node
    rec("", tree)
    return out


def _unflatten(flat, sep="/"):
    out = {}
    for k, v in flat.items():
        parts = k.split(sep)
        d = out
        for p in parts[:-1]:
            d = d.setdefault(p, {})
        d[parts[-1]] = v
    return out


def load_reference():
    ns = {"np": np, "dataclasses": dataclasses, "Enum": enum.Enum, "NormStats": NormStats,
          "traverse_util": types.SimpleNamespace(flatten_dict=_flatten, unflatten_dict=_unflatten),
          "DataTransformFn": object, "at": None, "T": None, "S": None, "Callable": None}
    exec(compile(defs(HELPERS, ["NormalizationType"]), HELPERS, "exec"), ns)
    exec(compile(defs(OP_T, ["flatten_dict", "unflatten_dict", "apply_tree", "_assert_quantile_stats"]), OP_T, "exec"), ns)
    exec(compile(defs(LAP_T, ["pad_to_dim", "Normalize", "Unnormalize", "PadStates"]), LAP_T, "exec"), ns)
    return ns


def cases():
    rng = np.random.default_rng(5)
    d = 7
    stats = dict(mean=rng.normal(size=d), std=rng.uniform(0.1, 2.0, size=d), q01=rng.normal(size=d) - 2.0,
                 q99=rng.normal(size=d) + 2.0, min=rng.normal(size=d) - 3.0, max=rng.normal(size=d) + 3.0)
    stats["q99"][3] = stats["q01"][3]  # zero-range dimension
    stats["max"][5] = stats["min"][5]
    data = dict(state=rng.normal(size=(d,)), actions=rng.normal(size=(10, d)) * 2.0, wide=rng.normal(size=(4, 32)),
                untouched=rng.normal(size=(3,)))
    return stats, data


def main():
    ref = load_reference()
    stats, data = cases()
    ns_ref = {k: NormStats(**stats) for k in ("state", "actions", "wide")}
    out = {f"stats/{k}": v for k, v in stats.items()}
    out.update({f"data/{k}": v for k, v in data.items()})
    for kind in ("normal", "bounds", "bounds_q99"):
        # (Normalize needs data no wider than the statistics; the wider tensor exercises Unnormalize's padding only)
        ns_norm = {k: v for k, v in ns_ref.items() if k != "wide"}
        norm = ref["Normalize"](ns_norm, kind)({k: v.copy() for k, v in data.items()})
        for k, v in norm.items():
            out[f"normalize/{kind}/{k}"] = np.asarray(v)
        un = ref["Unnormalize"](ns_ref, kind)({k: v.copy() for k, v in data.items()})
        for k, v in un.items():
            out[f"unnormalize/{kind}/{k}"] = np.asarray(v)
    out["padstates/short"] = ref["PadStates"](32)({"state": data["state"].copy()})["state"]
    out["padstates/long"] = ref["PadStates"](4)({"state": data["state"].copy()})["state"]
    np.savez_compressed(os.path.join(HERE, "reference_transforms.npz"), **out)
    print(sorted(out)[:6], len(out))


if __name__ == "__main__":
    main()
