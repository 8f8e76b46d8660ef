"""Assembling a serving policy from a config, a checkpoint and a tokenizer model — `create_trained_policy(_ar)` of
src/lap/policies/policy_config_adapter.py:86-160 with the transform groups of src/lap/training/config.py
(`RLDSDataConfig._create_data_transforms` :322-360, `ModelTransformFactory.__call__` :189-208) and the norm-stats file of
src/lap/shared/normalize_adapter.py (:26-59, `checkpoints.load_norm_stats` :477-497).

Order of the chain, as in the reference:
  inputs   repack.inputs, InjectDefaultPrompt, CoTInputs, Normalize, [InjectDefaultPrompt,] TokenizePromptAndReasoning,
           PadStatesAndActions
  outputs  DetokenizeReasoning, Unnormalize, CoTOutputs, repack.outputs          (standard strategy)
           DetokenizeReasoning, CoTOutputs(norm_stats, normalization_type), repack.outputs      (vla0 strategy)
What is passed in instead of looked up: the SentencePiece processor (a download in the reference) and the few `DataConfig`
fields the chain reads (defaults = RLDSDataConfig's, config.py:94-119).
"""
from __future__ import annotations

import dataclasses
import json
from pathlib import Path
from typing import Any

import numpy as np

from . import transforms as T
from .policy import ARPolicy, Policy
from .policy_io import CoTInputs, CoTOutputs


@dataclasses.dataclass(frozen=True)
class InjectDefaultPrompt:
    """OP/transforms.py:104-111."""
    prompt: str | None

    def __call__(self, data: dict) -> dict:
        if self.prompt is not None and "prompt" not in data:
            data["prompt"] = np.asarray(self.prompt)
        return data


def _pad_to_dim(x, target_dim: int, axis: int = -1):
    """OP/transforms.py:423-430: zero-pads, never truncates (unlike lap.transforms.pad_to_dim, which PadStates uses)."""
    x = np.asarray(x)
    if x.shape[axis] < target_dim:
        width = [(0, 0)] * x.ndim
        width[axis] = (0, target_dim - x.shape[axis])
        return np.pad(x, width)
    return x


@dataclasses.dataclass(frozen=True)
class PadStatesAndActions:
    """OP/transforms.py:327-337."""
    model_action_dim: int

    def __call__(self, data: dict) -> dict:
        data["state"] = _pad_to_dim(data["state"], self.model_action_dim, axis=-1)
        if "actions" in data:
            data["actions"] = _pad_to_dim(data["actions"], self.model_action_dim, axis=-1)
        return data


@dataclasses.dataclass(frozen=True)
class DataConfig:
    """The RLDSDataConfig fields the policy chain reads (src/lap/training/config.py:94-119, same names and defaults)."""
    wrist_image_dropout_prob: float = 0.1
    action_proprio_normalization_type: str = "bounds_q99"
    random_base_prob: float = 0.0
    random_mask_prob: float = 0.2
    use_rough_scale: bool = False
    language_action_format_name: str = "verbose_eef_with_rotation"
    transform_strategy: str = "standard"
    action_encoding: int = 1


def load_norm_stats(assets_dir) -> dict[str, T.NormStats]:
    """checkpoints.load_norm_stats (:477-497) + normalize_adapter.deserialize_json (:32-41): exactly one sub-directory of
    `assets_dir` holds `norm_stats.json` = {"norm_stats": {key: {mean, std, q01, q99[, min, max, ...]}}}; the key
    `state_eef_pose` is read as `state`."""
    assets_dir = Path(assets_dir)
    dirs = [p for p in assets_dir.iterdir() if p.is_dir() and (p / "norm_stats.json").exists()]
    assert len(dirs) == 1, f"Expected exactly one norm stats directory in {assets_dir}, but found {len(dirs)}: {[p.name for p in dirs]}"
    raw = json.loads((dirs[0] / "norm_stats.json").read_text())["norm_stats"]
    arr = lambda v: None if v is None else np.asarray(v)
    return {k.replace("state_eef_pose", "state"): T.NormStats(**{f: arr(v.get(f)) for f in ("mean", "std", "q01", "q99", "min", "max")})
            for k, v in raw.items()}


def policy_transforms(model_config, tokenizer, norm_stats, *, data: DataConfig = DataConfig(), default_prompt: str | None = None,
                      repack_inputs=(), repack_outputs=(), include_outputs: bool = True):
    """-> (input transforms, output transforms) in the reference's order."""
    nt = data.action_proprio_normalization_type
    inputs = [
        *repack_inputs,
        InjectDefaultPrompt(default_prompt),
        CoTInputs(action_dim=model_config.action_dim, model_type="lap", wrist_image_dropout_prob=data.wrist_image_dropout_prob,
                  action_encoding=data.action_encoding, language_action_format=data.language_action_format_name,
                  random_mask_prob=data.random_mask_prob, random_base_prob=data.random_base_prob,
                  use_rough_scale=data.use_rough_scale, transform_strategy=data.transform_strategy,
                  enable_langact_training=model_config.enable_langact_training),
        T.Normalize(norm_stats, normalization_type=nt),
        InjectDefaultPrompt(None),  # ModelTransformFactory.default_prompt (config.py:197)
        T.TokenizePromptAndReasoning(tokenizer, discrete_state_input=model_config.discrete_state_input,
                                     verbose_mode=model_config.verbose_mode, state_dropout=model_config.state_dropout),
        PadStatesAndActions(model_config.action_dim),
    ]
    model_outputs = [T.DetokenizeReasoning(tokenizer)] if include_outputs else []
    cot_out = CoTOutputs(language_action_format=data.language_action_format_name, transform_strategy=data.transform_strategy)
    if data.transform_strategy == "vla0":
        outputs = [*model_outputs, dataclasses.replace(cot_out, norm_stats=norm_stats, normalization_type=str(nt)), *repack_outputs]
    else:
        outputs = [*model_outputs, T.Unnormalize(norm_stats, normalization_type=nt), cot_out, *repack_outputs]
    return inputs, outputs


def create_trained_policy(train_config, checkpoint_dir, *, sp_processor, model=None, norm_stats=None, data: DataConfig = DataConfig(),
                          default_prompt: str | None = None, sample_kwargs: dict[str, Any] | None = None, repack_inputs=(),
                          repack_outputs=(), step: int | None = None) -> Policy:
    """policy_config_adapter.py:86-155.  `checkpoint_dir` holds step directories written by `lap_b200.checkpoint` (the served
    weights are the `params` item: EMA weights when the run used EMA) and, unless `norm_stats` is given, `assets/<id>/
    norm_stats.json`.  `model`: an already constructed `LAP` to load into (default: a new one — needs the CUDA engine)."""
    from . import checkpoint as _checkpoints
    checkpoint_dir = Path(checkpoint_dir)
    if model is None:
        from .model import LAP
        model = LAP(train_config.model, init=False)
    _checkpoints.load_served_params(checkpoint_dir, model, step)
    if norm_stats is None:
        norm_stats = load_norm_stats(checkpoint_dir / "assets")
    tokenizer = train_config.model.make_tokenizer(sp_processor)
    inputs, outputs = policy_transforms(train_config.model, tokenizer, norm_stats, data=data, default_prompt=default_prompt,
                                        repack_inputs=repack_inputs, repack_outputs=repack_outputs)
    return Policy(model, transforms=inputs, output_transforms=outputs, sample_kwargs=sample_kwargs)


def create_trained_policy_ar(*args, sample_kwargs: dict | None = None, **kwargs) -> ARPolicy:
    """policy_config_adapter.py:157-160."""
    return ARPolicy(create_trained_policy(*args, **kwargs), sample_kwargs=sample_kwargs)
