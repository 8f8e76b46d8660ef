"""Host-side normalisation transforms of `Policy.infer` (first slice) — numpy only.

Reference: src/lap/transforms.py:150-289 (`Normalize`, `Unnormalize`, `PadStates`), :554-562 (`pad_to_dim`);
third_party/openpi/src/openpi/transforms.py:340-347,404-420,455-460 (`flatten_dict`, `apply_tree`,
`_assert_quantile_stats`); third_party/openpi/src/openpi/shared/normalize.py:10-14 (`NormStats`);
src/lap/datasets/utils/helpers.py:32-37 (`NormalizationType`).  Same class names, fields, defaults and error behaviour, so a
`Policy(model, transforms=[..., Normalize(stats, "bounds_q99")], output_transforms=[Unnormalize(stats, "bounds_q99"), ...])`
reads like the reference's `policy_config` wiring.  `TokenizePromptAndReasoning` / `DetokenizeReasoning` /
`SafeRepackTransform` (src/lap/transforms.py:26-147) wrap `lap_b200.tokenizer.CoTTokenizer`.  Checked against the reference classes executed from source
(tests/golden/make_reference_transforms_golden.py -> tests/golden/reference_transforms.npz).
"""
from __future__ import annotations

import dataclasses
import enum
from collections.abc import Callable
from typing import Any

import numpy as np


class NormalizationType(str, enum.Enum):
    NORMAL = "normal"          # mean 0, std 1
    BOUNDS = "bounds"          # [min, max] -> [-1, 1], clipped
    BOUNDS_Q99 = "bounds_q99"  # [q01, q99] -> [-1, 1], not clipped


@dataclasses.dataclass
class NormStats:
    mean: np.ndarray
    std: np.ndarray
    q01: np.ndarray | None = None
    q99: np.ndarray | None = None
    min: np.ndarray | None = None
    max: np.ndarray | None = None


def flatten_dict(tree: dict, sep: str = "/") -> dict:
    out: dict[str, Any] = {}

    def rec(prefix: str, node: Any) -> None:
        if isinstance(node, dict) and node:
            for k, v in node.items():
                rec(f"{prefix}{sep}{k}" if prefix else str(k), v)
        else:
            out[prefix] = node

    rec("", tree)
    return out


def unflatten_dict(flat: dict, sep: str = "/") -> dict:
    out: dict[str, Any] = {}
    for k, v in flat.items():
        parts = k.split(sep)
        d = out
        for p in parts[:-1]:
            d = d.setdefault(p, {})
        d[parts[-1]] = v
    return out


def apply_tree(tree: dict, selector: dict, fn: Callable[[Any, Any], Any], *, strict: bool = False) -> dict:
    """OP/transforms.py:404-420."""
    tree, selector = flatten_dict(tree), flatten_dict(selector)
    if strict:
        for k in selector:
            if k not in tree:
                raise ValueError(f"Selector key {k} not found in tree")
    return unflatten_dict({k: (fn(v, selector[k]) if k in selector else v) for k, v in tree.items()})


def pad_to_dim(x: np.ndarray, target_dim: int, axis: int = -1, value: float = 0.0) -> np.ndarray:
    """transforms.py:554-562 (pads along `axis`, truncates along the LAST axis — as the reference does)."""
    current_dim = x.shape[axis]
    if current_dim < target_dim:
        pad_width = [(0, 0)] * len(x.shape)
        pad_width[axis] = (0, target_dim - current_dim)
        return np.pad(x, pad_width, constant_values=value)
    return x[..., :target_dim]


def _assert_quantile_stats(norm_stats: dict) -> None:
    for k, v in flatten_dict(norm_stats).items():
        if v.q01 is None or v.q99 is None:
            raise ValueError(f"quantile stats must be provided if use_quantile_norm is True. Key {k} is missing q01 or q99.")


def _resolve(t: NormalizationType | str) -> NormalizationType:
    return NormalizationType(t) if isinstance(t, str) else t


@dataclasses.dataclass(frozen=True)
class Normalize:
    """transforms.py:150-218."""
    norm_stats: dict | None
    normalization_type: NormalizationType | str = NormalizationType.NORMAL
    strict: bool = False

    def __post_init__(self):
        if self.norm_stats is not None and _resolve(self.normalization_type) == NormalizationType.BOUNDS_Q99:
            _assert_quantile_stats(self.norm_stats)

    def __call__(self, data: dict) -> dict:
        if self.norm_stats is None:
            return data
        fn = {NormalizationType.NORMAL: self._normalize, NormalizationType.BOUNDS: self._normalize_bounds,
              NormalizationType.BOUNDS_Q99: self._normalize_quantile}[_resolve(self.normalization_type)]
        return apply_tree(data, self.norm_stats, fn, strict=self.strict)

    @staticmethod
    def _normalize(x, stats: NormStats):
        mean, std = stats.mean[..., : x.shape[-1]], stats.std[..., : x.shape[-1]]
        return (x - mean) / (std + 1e-6)

    @staticmethod
    def _normalize_bounds(x, stats: NormStats):
        assert stats.min is not None and stats.max is not None
        lo, hi = stats.min[..., : x.shape[-1]], stats.max[..., : x.shape[-1]]
        scaled = np.clip(2.0 * (x - lo) / (hi - lo + 1e-8) - 1.0, -1.0, 1.0)
        zeros = np.equal(lo, hi)
        while zeros.ndim < x.ndim:
            zeros = zeros[None, ...]
        return np.where(zeros, 0.0, scaled)

    @staticmethod
    def _normalize_quantile(x, stats: NormStats):
        assert stats.q01 is not None and stats.q99 is not None
        q01, q99 = stats.q01[..., : x.shape[-1]], stats.q99[..., : x.shape[-1]]
        scaled = (x - q01) / (q99 - q01 + 1e-6) * 2.0 - 1.0
        zeros = np.equal(q01, q99)
        while zeros.ndim < x.ndim:
            zeros = zeros[None, ...]
        return np.where(zeros, 0.0, scaled)


@dataclasses.dataclass(frozen=True)
class Unnormalize:
    """transforms.py:220-278."""
    norm_stats: dict | None
    normalization_type: NormalizationType | str = NormalizationType.NORMAL

    def __post_init__(self):
        if self.norm_stats is not None and _resolve(self.normalization_type) == NormalizationType.BOUNDS_Q99:
            _assert_quantile_stats(self.norm_stats)

    def __call__(self, data: dict) -> dict:
        if self.norm_stats is None:
            return data
        fn = {NormalizationType.NORMAL: self._unnormalize, NormalizationType.BOUNDS: self._unnormalize_bounds,
              NormalizationType.BOUNDS_Q99: self._unnormalize_quantile}[_resolve(self.normalization_type)]
        return apply_tree(data, self.norm_stats, fn, strict=False)

    @staticmethod
    def _unnormalize(x, stats: NormStats):
        mean = pad_to_dim(stats.mean, x.shape[-1], axis=-1, value=0.0)
        std = pad_to_dim(stats.std, x.shape[-1], axis=-1, value=1.0)
        return x * (std + 1e-6) + mean

    @staticmethod
    def _unnormalize_bounds(x, stats: NormStats):
        assert stats.min is not None and stats.max is not None
        lo = pad_to_dim(stats.min, x.shape[-1], axis=-1, value=-1.0)
        hi = pad_to_dim(stats.max, x.shape[-1], axis=-1, value=1.0)
        return (x + 1.0) / 2.0 * (hi - lo + 1e-8) + lo

    @staticmethod
    def _unnormalize_quantile(x, stats: NormStats):
        assert stats.q01 is not None and stats.q99 is not None
        q01, q99 = stats.q01, stats.q99
        if (dim := q01.shape[-1]) < x.shape[-1]:
            return np.concatenate([(x[..., :dim] + 1.0) / 2.0 * (q99 - q01 + 1e-6) + q01, x[..., dim:]], axis=-1)
        return (x + 1.0) / 2.0 * (q99 - q01 + 1e-6) + q01


@dataclasses.dataclass(frozen=True)
class PadStates:
    """transforms.py:280-289: zero-pads (or truncates) `state` to the model action dimension."""
    model_action_dim: int

    def __call__(self, data: dict) -> dict:
        data["state"] = pad_to_dim(data["state"], self.model_action_dim, axis=-1)
        return data


@dataclasses.dataclass(frozen=True)
class TokenizePromptAndReasoning:
    """src/lap/transforms.py:26-110: pops `prompt` / `language_actions` / `dataset_name` / `frame_description` /
    `time_horizon_seconds` from the sample and adds the token / mask fields `CoTObservation.from_dict` reads.
    `language_actions` absent (inference) -> `tokenized_langact_mask` is None."""
    tokenizer: Any
    discrete_state_input: bool = False
    dataset_name_pad_len: int = 100
    verbose_mode: bool = False
    state_dropout: float = 0.0

    def __call__(self, data: dict) -> dict:
        prompt = data.pop("prompt", None)
        if prompt is None:
            raise ValueError("Prompt is required")
        if not isinstance(prompt, str):
            prompt = prompt.item()
        state = None
        if self.discrete_state_input:
            state = data.get("state", None)
            if state is None:
                raise ValueError("State is required.")
        language_actions = data.pop("language_actions", None)
        dataset_name = data.pop("dataset_name", None)
        frame_description = data.pop("frame_description", "robot base frame")
        sp = self.tokenizer._tokenizer
        name_ids = sp.encode(dataset_name) if dataset_name is not None else []
        # left padded; a name longer than the pad length is kept whole, as in the reference ([pad] * negative == [])
        name_ids = [sp.pad_id()] * (self.dataset_name_pad_len - len(name_ids)) + name_ids
        is_vqa_sample, is_prediction_sample = data["is_vqa_sample"], data["is_prediction_sample"]
        time_horizon_seconds = data.pop("time_horizon_seconds", None)
        tokens, pad_mask, reasoning_mask, number_mask, direction_mask, token_loss_mask = self.tokenizer.tokenize(
            prompt, language_actions, state, is_vqa_sample=is_vqa_sample, is_prediction_sample=is_prediction_sample,
            time_horizon_seconds=time_horizon_seconds, frame_description=frame_description,
            state_dropout=self.state_dropout)
        out = {**data, "tokenized_prompt": tokens, "tokenized_prompt_mask": pad_mask,
               "tokenized_langact_mask": reasoning_mask, "token_loss_mask": token_loss_mask,
               "tokenized_dataset_name": np.asarray(name_ids, dtype=np.int32)}
        if self.verbose_mode:
            # (np.logical_or(None, None) is None: inference samples carry no critical mask)
            out.update(critical_token_mask=np.logical_or(number_mask, direction_mask), number_token_mask=number_mask,
                       direction_token_mask=direction_mask)
        return out


@dataclasses.dataclass(frozen=True)
class DetokenizeReasoning:
    """src/lap/transforms.py:112-120: `tokens` (the output of `sample_tokens`) -> `reasoning` text."""
    tokenizer: Any

    def __call__(self, data: dict) -> dict:
        if "tokens" in data:
            return {**data, "reasoning": self.tokenizer.decode(np.asarray(data["tokens"]).squeeze().astype(np.int32))}
        return data


@dataclasses.dataclass(frozen=True)
class SafeRepackTransform:
    """src/lap/transforms.py:123-147: `structure` maps output paths to a source path or a list of fall-back source paths
    ('/'-joined); missing sources are skipped unless `strict`."""
    structure: Any
    strict: bool = False

    def __call__(self, data: dict) -> dict:
        flat = flatten_dict(data)
        out, missing = {}, []
        for key, spec in flatten_dict(self.structure).items():
            candidates = spec if isinstance(spec, (list, tuple)) else [spec]
            hit = next((c for c in candidates if c in flat), None)
            if hit is None:
                missing.append((key, tuple(candidates)))
            else:
                out[key] = flat[hit]
        if self.strict and missing:
            raise KeyError(f"Missing source paths: {missing}")
        return unflatten_dict(out)


@dataclasses.dataclass(frozen=True)
class ResizeImages:
    """third_party/openpi/src/openpi/transforms.py:184-191, with the JAX path's `resize_with_pad` (lap_b200.image_tools) — put
    it in front of the tokenizer transform when a client sends images that are not already height x width."""
    height: int
    width: int

    def __call__(self, data: dict) -> dict:
        from .image_tools import resize_with_pad
        data["image"] = {k: resize_with_pad(np.asarray(v), self.height, self.width) for k, v in data["image"].items()}
        return data
