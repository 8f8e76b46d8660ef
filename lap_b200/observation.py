"""Observation structs of the hot-path boundary (same field names as the reference).

  Observation ....... third_party/openpi/src/openpi/models/model.py:81-136
  CoTObservation .... src/lap/models/model_adapter.py:37-80
Arrays are numpy (host) or torch tensors; the engine stages them to the device itself and never aliases caller
memory after a call returns (SURVEY §8b "data ownership").
"""
from __future__ import annotations

import dataclasses
from typing import Any

import numpy as np


@dataclasses.dataclass
class Observation:
    images: dict[str, Any]  # [B,H,W,3] float32 in [-1,1] (or uint8, converted on the device)
    image_masks: dict[str, Any]  # [B] bool
    state: Any  # [B, s] float32 (unused when pi05: the state is discretised into the prompt)
    tokenized_prompt: Any = None  # [B, L] int32
    tokenized_prompt_mask: Any = None  # [B, L] bool
    token_ar_mask: Any = None
    token_loss_mask: Any = None

    @classmethod
    def from_dict(cls, data: dict) -> "Observation":
        """model.py:109-129.  uint8 images are kept uint8 here: the u8/255*2-1 conversion (model.py:116-118) is
        fused into the device-side patchify kernel instead of being an eager pass over the batch."""
        if ("tokenized_prompt" in data) != ("tokenized_prompt_mask" in data):
            raise ValueError("tokenized_prompt and tokenized_prompt_mask must be provided together.")
        return cls(
            images=dict(data["image"]),
            image_masks=dict(data["image_mask"]),
            state=data["state"],
            tokenized_prompt=data.get("tokenized_prompt"),
            tokenized_prompt_mask=data.get("tokenized_prompt_mask"),
            token_ar_mask=data.get("token_ar_mask"),
            token_loss_mask=data.get("token_loss_mask"),
        )

    def to_dict(self) -> dict:
        d = dataclasses.asdict(self)
        d["image"] = d.pop("images")
        d["image_mask"] = d.pop("image_masks")
        return d


@dataclasses.dataclass
class CoTObservation(Observation):
    tokenized_langact_mask: Any = None
    critical_token_mask: Any = None
    number_token_mask: Any = None
    direction_token_mask: Any = None
    sample_mask: Any = None
    tokenized_dataset_name: Any = None
    is_vqa_sample: Any = None
    is_prediction_sample: Any = None
    vqa_dataset_id: Any = None

    @classmethod
    def from_dict(cls, data: dict) -> "CoTObservation":
        """model_adapter.py:51-80 (flat keys or `extras/cot` namespace)."""
        base = Observation.from_dict(data)
        cot = data.get("extras", {}).get("cot", {}) if isinstance(data.get("extras"), dict) else {}

        def getk(k):
            return data.get(k, cot.get(k))

        return cls(
            **{f.name: getattr(base, f.name) for f in dataclasses.fields(Observation)},
            tokenized_langact_mask=getk("tokenized_langact_mask"),
            critical_token_mask=getk("critical_token_mask"),
            number_token_mask=getk("number_token_mask"),
            direction_token_mask=getk("direction_token_mask"),
            sample_mask=getk("sample_mask"),
            tokenized_dataset_name=getk("tokenized_dataset_name"),
            is_vqa_sample=getk("is_vqa_sample"),
            is_prediction_sample=getk("is_prediction_sample"),
            vqa_dataset_id=getk("vqa_dataset_id"),
        )


def to_numpy(x):
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        return x
    if hasattr(x, "detach"):
        return x.detach().cpu().numpy()
    return np.asarray(x)
