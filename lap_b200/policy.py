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
    if isinstance(x, dict):
        return {k: _tree_map(f, v) for k, v in x.items()}
    return f(x)


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
