"""`Policy.infer` surface — third_party/openpi/src/openpi/policies/policy.py:24-110.

Same constructor/`infer` contract as the reference `Policy`: host-side input transforms -> batch of 1 ->
`Observation.from_dict` -> `model.sample_actions` -> unbatch -> output transforms -> `policy_timing.infer_ms`.
The transforms themselves (tokenizer, normalisation, ...) are caller-supplied callables: they are outside the hot path
 and work unchanged on top of this class.  Single-threaded like the reference (the websocket server calls
`infer` synchronously, websocket_policy_server.py:61); not re-entrant.
"""
from __future__ import annotations

import time
from collections.abc import Callable, Sequence
from typing import Any

import numpy as np
import torch

from .observation import Observation


def _compose(fns: Sequence[Callable[[dict], dict]]) -> Callable[[dict], dict]:
    def run(d: dict) -> dict:
        for f in fns:
            d = f(d)
        return d

    return run


def _tree_map(f, x):
    """jax.tree.map over nested dicts; None is an empty subtree (stays None), as in JAX."""
    if isinstance(x, dict):
        return {k: _tree_map(f, v) for k, v in x.items()}
    return None if x is None else f(x)


class Policy:
    def __init__(self, model, *, rng: int | None = None, transforms: Sequence[Callable] = (),
                 output_transforms: Sequence[Callable] = (), sample_kwargs: dict[str, Any] | None = None,
                 metadata: dict[str, Any] | None = None):
        self._model = model
        self._input_transform = _compose(transforms)
        self._output_transform = _compose(output_transforms)
        self._sample_kwargs = sample_kwargs or {}
        self._metadata = metadata or {}
        self._rng = 0 if rng is None else int(rng)

    def infer(self, obs: dict, *, noise: np.ndarray | None = None) -> dict:
        inputs = _tree_map(lambda x: x, obs)  # copy: transforms may modify in place (policy.py:70)
        inputs = self._input_transform(inputs)
        inputs = _tree_map(lambda x: np.asarray(x)[np.newaxis, ...], inputs)  # batch of 1 (policy.py:74)
        self._rng += 1
        sample_kwargs = dict(self._sample_kwargs)
        if noise is not None:
            noise = np.asarray(noise)
            if noise.ndim == 2:
                noise = noise[None, ...]
            sample_kwargs["noise"] = noise
        observation = Observation.from_dict(inputs)
        start = time.monotonic()
        actions = self._model.sample_actions(self._rng, observation, **sample_kwargs)
        actions = actions.cpu().numpy() if isinstance(actions, torch.Tensor) else np.asarray(actions)
        model_time = time.monotonic() - start
        outputs = {"state": np.asarray(inputs["state"])[0], "actions": actions[0]}
        outputs = self._output_transform(outputs)
        outputs["policy_timing"] = {"infer_ms": model_time * 1000}
        return outputs

    def reset(self) -> None:  # openpi_client.BasePolicy.reset
        pass

    @property
    def metadata(self) -> dict[str, Any]:
        return self._metadata


class ARPolicy:
    """src/lap/policies/policy_adapter.py:13-61: the reasoning-decoding policy.  Wraps a `Policy`; `infer` runs the input
    transforms, `model.sample_tokens` (greedy autoregressive decode of the language-action text) and the output transforms
    (`DetokenizeReasoning` -> `CoTOutputs`), which see `{"state", "tokens", "raw_state"}` - batched, like the reference."""

    def __init__(self, base: Policy, *, sample_kwargs: dict[str, Any] | None = None):
        assert hasattr(base._model, "sample_tokens"), "Model must have a sample_tokens method"
        self._base = base
        self._ar_kwargs = sample_kwargs or {}

    def __getattr__(self, name: str):
        return getattr(self._base, name)

    def infer_reasoning(self, obs: dict) -> dict:
        from .observation import CoTObservation
        inputs = _tree_map(lambda x: x, obs)
        raw_state = np.array(inputs["observation"]["state"], copy=True)
        inputs = self._base._input_transform(inputs)
        inputs = _tree_map(lambda x: np.asarray(x)[np.newaxis, ...], inputs)
        self._base._rng += 1
        start = time.monotonic()
        tokens = self._base._model.sample_tokens(self._base._rng, CoTObservation.from_dict(inputs), **self._ar_kwargs)
        tokens = tokens.cpu().numpy() if isinstance(tokens, torch.Tensor) else np.asarray(tokens)
        model_time = time.monotonic() - start
        outputs = self._base._output_transform({"state": inputs["state"], "tokens": tokens, "raw_state": raw_state})
        outputs["policy_timing"] = {"infer_ms": model_time * 1000}
        return outputs

    def infer(self, obs: dict, *, noise: np.ndarray | None = None) -> dict:
        return self.infer_reasoning(obs)

    def vqa_infer(self, obs: dict) -> dict:
        return self.infer_reasoning(obs)
