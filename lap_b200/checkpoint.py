"""Checkpoint save / restore of the engine's train state (the "write it back" half) in the REFERENCE layout.

What mirrors the reference (src/lap/training/checkpoints.py):
  * one directory per step; written to `<step>.tmp-<pid>` and renamed when complete (an interrupted save never looks valid);
  * the `_split_params` / `_merge_params` convention (:529-547): item `params` holds the weights to SERVE — the EMA weights when
    EMA is enabled, the raw weights otherwise — and item `train_state` holds the rest (step, raw params if EMA is on, Adam mu/nu);
  * `latest_step`; retention = `keep` newest checkpoints plus every `keep_period`-th step (the reference: max_to_keep=1).
Every tensor is stored under its reference parameter-tree path ('/'-joined, no `value` leaf), in the reference's shapes and
fp32, so `params.safetensors` of a checkpoint IS a reference-layout `params` tree (`LAP.load_params(load_tree(...))`).
What does not: the container is safetensors + a JSON manifest, not Orbax/OCDBT (orbax is not installable offline); reading a
released Orbax checkpoint needs `orbax` once to dump the tree to numpy, after which `LAP.load_params` takes it as is.
"""
from __future__ import annotations

import json
import os
import shutil
from pathlib import Path

import torch

from . import params as P

FORMAT = "lap_b200.checkpoint.v1"


def save_tree(path, tree: dict[str, torch.Tensor]) -> None:
    from safetensors.torch import save_file
    save_file({k: v.detach().to("cpu", torch.float32).contiguous() for k, v in tree.items()}, str(path))


def load_tree(path) -> dict[str, torch.Tensor]:
    from safetensors.torch import load_file
    return load_file(str(path))


def _steps(directory: Path) -> list[int]:
    if not directory.is_dir():
        return []
    return sorted(int(p.name) for p in directory.iterdir() if p.is_dir() and p.name.isdigit() and (p / "meta.json").exists())


def latest_step(directory) -> int | None:
    s = _steps(Path(directory))
    return s[-1] if s else None


def _model_signature(cfg) -> dict:
    shapes = P.reference_shapes(cfg)
    return {"n_tensors": len(shapes), "n_params": int(sum(int(torch.Size(s).numel()) for s in shapes.values()))}


def save_train_state(directory, state, step: int | None = None, *, keep: int | None = None,
                     keep_period: int | None = None) -> Path:
    """Write checkpoint `<directory>/<step>/`.  Call on rank 0 only (the state is replicated across data-parallel ranks).
    Retention as in the reference's CheckpointManagerOptions(max_to_keep, keep_period) (checkpoints.py:58-63, where
    max_to_keep = 1): only the `keep` newest checkpoints stay, except that steps divisible by `keep_period` are never removed."""
    directory = Path(directory)
    step = int(state.step if step is None else step)
    final = directory / str(step)
    tmp = directory / f"{step}.tmp-{os.getpid()}"
    if tmp.exists():
        shutil.rmtree(tmp)
    (tmp / "train_state").mkdir(parents=True)
    model = state.model
    has_ema = state.ema_params is not None
    save_tree(tmp / "params.safetensors", model.params_reference(state.ema_params if has_ema else model.P))
    if has_ema:
        save_tree(tmp / "train_state" / "params.safetensors", model.params_reference(model.P))
    save_tree(tmp / "train_state" / "mu.safetensors", model.params_reference(state.mu))
    save_tree(tmp / "train_state" / "nu.safetensors", model.params_reference(state.nu))
    # `step` names the directory (the loop step that triggered the save); `state_step` is TrainState.step, which the train step
    # has already advanced - resuming continues from it, as scripts/train.py does with the restored state (:528-530)
    meta = {"format": FORMAT, "step": step, "state_step": int(state.step), "has_ema": has_ema, "ema_decay": state.ema_decay,
            "model": _model_signature(model.cfg)}
    (tmp / "meta.json").write_text(json.dumps(meta, indent=1))
    if final.exists():
        shutil.rmtree(final)
    os.replace(tmp, final)
    if keep is not None:
        for s in _steps(directory)[:-keep]:
            if keep_period and s % keep_period == 0:
                continue
            shutil.rmtree(directory / str(s))
    return final


def _into_flat(model, flat: torch.Tensor, tree: dict[str, torch.Tensor]) -> None:
    want = P.reference_shapes(model.cfg)
    missing = [k for k in want if k not in tree]
    if missing:
        raise ValueError(f"checkpoint is missing {len(missing)} tensors, e.g. {missing[:3]}")
    for k, s in want.items():
        if tuple(tree[k].shape) != tuple(s):
            raise ValueError(f"shape mismatch for {k}: got {tuple(tree[k].shape)}, expected {tuple(s)}")
    eng = P.reference_to_engine(model.cfg, {k: tree[k].to(torch.float32) for k in want})
    for name in model.layout.shapes:
        model.layout.view(flat, name).copy_(eng[name])


def restore_train_state(directory, state, step: int | None = None) -> int:
    """Load checkpoint `step` (default: the latest) into `state` IN PLACE (params, Adam moments, EMA, step) and refresh the
    bf16 compute copy.  Returns the restored `TrainState.step`.  Raises FileNotFoundError / ValueError on a missing or mismatching one."""
    directory = Path(directory)
    step = latest_step(directory) if step is None else int(step)
    if step is None or not (directory / str(step) / "meta.json").exists():
        raise FileNotFoundError(f"no checkpoint{'' if step is None else f' for step {step}'} under {directory}")
    d = directory / str(step)
    meta = json.loads((d / "meta.json").read_text())
    if meta.get("format") != FORMAT:
        raise ValueError(f"{d}: unknown checkpoint format {meta.get('format')!r}")
    if meta["model"] != _model_signature(state.model.cfg):
        raise ValueError(f"{d}: checkpoint is for a different model ({meta['model']} vs {_model_signature(state.model.cfg)})")
    model = state.model
    served = load_tree(d / "params.safetensors")
    if meta["has_ema"]:  # _merge_params: `params` held the EMA weights
        if state.ema_params is None:
            state.ema_params = torch.empty_like(model.P)
        _into_flat(model, state.ema_params, served)
        _into_flat(model, model.P, load_tree(d / "train_state" / "params.safetensors"))
    else:
        _into_flat(model, model.P, served)
        state.ema_params = None
    _into_flat(model, state.mu, load_tree(d / "train_state" / "mu.safetensors"))
    _into_flat(model, state.nu, load_tree(d / "train_state" / "nu.safetensors"))
    state.ema_decay = meta["ema_decay"]
    state.step = int(meta.get("state_step", meta["step"]))
    model.refresh_compute_copy()
    return state.step


def load_served_params(directory, model, step: int | None = None) -> int:
    """Inference-side restore: only the `params` item (EMA weights when the run used EMA), as `serve_policy.py` does."""
    directory = Path(directory)
    step = latest_step(directory) if step is None else int(step)
    if step is None:
        raise FileNotFoundError(f"no checkpoint under {directory}")
    d = directory / str(step)
    if (d / "params.safetensors").exists():
        model.load_params(load_tree(d / "params.safetensors"))
    elif (d / "params").is_dir():
        # a checkpoint written by the REFERENCE (Orbax `params` item, third_party/openpi/src/openpi/models/model.py:286-332)
        from . import orbax_io
        model.load_params(orbax_io.read_params(d / "params"))
    else:
        raise FileNotFoundError(f"{d} holds neither params.safetensors nor an Orbax `params` item")
    return step


def export_orbax_params(directory, model, step: int, *, ema_flat: torch.Tensor | None = None, value_suffix: bool = True) -> Path:
    """Write `<directory>/<step>/params` as an Orbax item in the plain-directory layout (orbax_io.write_params) so that the
    reference's `restore_params` / `serve_policy.py` can load weights trained here.  `ema_flat` = the EMA buffer when the
    run used EMA (`_split_params`, checkpoints.py:529-538: the served `params` item holds the EMA weights)."""
    from . import orbax_io
    out = Path(directory) / str(int(step)) / "params"
    tree = model.params_reference(ema_flat if ema_flat is not None else None)
    orbax_io.write_params(out, {k: v.numpy() for k, v in tree.items()}, value_suffix=value_suffix)
    return out
