"""Request / sample adapters either side of the model (fourth slice): `CoTInputs` turns a raw robot sample or
a serving request into the dict `TokenizePromptAndReasoning` and `CoTObservation.from_dict` consume; `CoTOutputs` turns the
decoded reasoning text back into an action.

Reference: src/lap/policies/transforms/input_transforms.py:24-249 (`CoTInputs`), image_handler.py (`ImageHandler`),
image_utils.py (`parse_image`), text_utils.py (`TextParser`), action_processor.py (`ActionProcessor`),
sample_handlers.py:44-70,328-457 (VQA and robot samples), output_transforms.py (`CoTOutputs`).
The reference spreads this over handler / strategy / processor classes; here it is two dataclasses and a few functions.
Diverse prediction questions (`question_types.py`, `PredictionSampleHandler`) are a training-data augmentation outside the
path: `enable_diverse_questions=True` raises NotImplementedError.  Randomness (wrist dropout, zero-image unmasking, random
base frame) draws from `np.random` / `random` in the same order as the reference.
Checked against the reference classes executed from source (tests/golden/make_reference_langaction_golden.py).
"""
from __future__ import annotations

import dataclasses
import random

import numpy as np

from . import lang_actions as LA
from .transforms import pad_to_dim

IMAGE_KEYS = ("base_0_rgb", "left_wrist_0_rgb")  # src/lap/models/model_adapter.py:18-22
_MODEL_TYPES = ("lap", "lap_fast", "pi0_fast")   # input_transforms.py:143-147


def parse_image(image):
    """image_utils.py:7-17 - float [0,1] -> uint8, CHW / TCHW -> HWC / THWC."""
    if image is None:
        return None
    image = np.asarray(image)
    if np.issubdtype(image.dtype, np.floating):
        image = (255 * image).astype(np.uint8)
    if image.ndim == 3 and image.shape[0] == 3:
        image = np.transpose(image, (1, 2, 0))
    if image.ndim == 4 and image.shape[1] == 3:
        image = np.transpose(image, (0, 2, 3, 1))
    return image


def decode_text(value, default: str = "") -> str:
    """text_utils.py:8-22."""
    if isinstance(value, bytes):
        return value.decode("utf-8")
    return value if isinstance(value, str) else default


def parse_prompt(data: dict) -> str:
    """text_utils.py:37-60 (r1_lite prompts carry an "...@" prefix)."""
    prompt = data.get("prompt")
    assert prompt is not None, "Prompt missing from data"
    text = decode_text(prompt)
    return text.split("@")[-1] if "r1_lite" in decode_text(data.get("dataset_name")) else text


def image_mask(image, random_mask_prob: float = 0.0):
    """image_handler.py:24-40 - an all-zero image is masked out, except with probability `random_mask_prob`."""
    if np.all(image == 0.0):
        return np.True_ if random_mask_prob > 0.0 and np.random.rand() < random_mask_prob else np.False_
    return np.True_


@dataclasses.dataclass(frozen=True)
class CoTInputs:
    action_dim: int
    language_action_format: LA.LanguageActionFormat | str | None = LA.VERBOSE_EEF_WITH_ROTATION_FORMAT
    wrist_image_dropout_prob: float = 0.0
    model_type: str = "lap"
    action_encoding: int = 1                       # ActionEncoding.EEF_POS (datasets/utils/helpers.py:23-28)
    enable_langact_training: bool = True
    use_rough_scale: bool = False
    transform_strategy: str = "standard"
    random_base_prob: float = 0.0
    random_mask_prob: float = 0.0
    enable_diverse_questions: bool = False

    def __post_init__(self):
        if isinstance(self.language_action_format, str):
            object.__setattr__(self, "language_action_format", LA.get_language_action_format(self.language_action_format))
        if self.enable_diverse_questions:
            raise NotImplementedError("diverse prediction questions (question_types.py) are outside the hot path")

    # image_handler.py:42-119,146-166
    def _images(self, data, is_prediction, pred_use_primary, is_vqa):
        obs = data.get("observation", {})
        raw = data["observation"].get(IMAGE_KEYS[0])
        base = None if isinstance(raw, (str, bytes)) and len(raw) == 0 else parse_image(raw)
        if base is None:
            base = np.zeros((224, 224, 3), dtype=np.uint8)   # masked out below
        images, masks = [], []

        def add(img, p=0.0):
            masks.append(image_mask(img, p))
            images.append(img)

        if not is_prediction:
            add(base)
            for key in IMAGE_KEYS[1:]:
                if key not in obs:
                    wrist = np.zeros_like(base)
                else:
                    wrist = parse_image(obs[key])
                    if (not is_vqa and self.wrist_image_dropout_prob > 0.0
                            and np.random.rand() < float(self.wrist_image_dropout_prob)):
                        wrist = np.zeros_like(base)
                add(wrist, 0.0 if is_vqa else self.random_mask_prob)
        else:
            keys = IMAGE_KEYS[1:] if pred_use_primary else IMAGE_KEYS
            if pred_use_primary:
                add(base)
            for key in keys:
                add(parse_image(obs[key]) if key in obs else np.zeros_like(base))
        return images, masks

    # action_processor.py:31-131
    def _summarize(self, data, initial_state, dataset_name, rotation_applied):
        fmt = self.language_action_format
        acts = data["language_actions"]
        use_eef = fmt.use_eef_frame and initial_state is not None
        if self.random_base_prob > 0.0:
            use_eef = use_eef and data.get("has_wrist_image", False) and random.random() < (1 - self.random_base_prob)
        if use_eef:
            acts = LA.transform_actions_to_eef_frame(acts, initial_state, dataset_name, rotation_applied)
        if data.get("is_bimanual", False):
            text = LA.summarize_bimanual_numeric_actions(acts, fmt.get_sum_decimal(), fmt.include_rotation)
        elif data.get("is_navigation", False):
            text = LA.summarize_numeric_actions(acts, "nearest_10", include_rotation=True, rotation_precision=10)
        else:
            text = LA.summarize_numeric_actions(acts, sum_decimal=fmt.get_sum_decimal(), include_rotation=fmt.include_rotation)
        return text, ("end-effector frame" if use_eef else "robot base frame")

    def __call__(self, data: dict) -> dict:
        assert self.model_type in _MODEL_TYPES
        assert "observation" in data
        dataset_name = decode_text(data.get("dataset_name"))
        is_prediction = data.get("is_prediction_sample", False)
        is_vqa = data.get("is_vqa_sample", False)
        images, masks = self._images(data, is_prediction, data.get("pred_use_primary", False), is_vqa)
        if self.model_type == "lap_fast":
            masks = [np.True_ for _ in masks]
        inputs = {"state": data["observation"]["state"], "image": dict(zip(IMAGE_KEYS, images, strict=True)),
                  "image_mask": dict(zip(IMAGE_KEYS, masks, strict=True)), "prompt": parse_prompt(data),
                  "is_prediction_sample": is_prediction}
        if dataset_name:
            inputs["dataset_name"] = dataset_name
        if "frame_description" in data:
            inputs["frame_description"] = decode_text(data["frame_description"], default="robot base frame")
        if "actions" in data:
            inputs["actions"] = np.array(pad_to_dim(data["actions"], self.action_dim))
        rotation_applied = data.get("rotation_applied", False)
        inputs["is_vqa_sample"] = is_vqa
        inputs["time_horizon_seconds"] = data.get("time_horizon_seconds")
        inputs["vqa_dataset_id"] = data.get("vqa_dataset_id", 0)

        if is_vqa:  # sample_handlers.py:52-70: the caption is the target text, never filtered
            caption = data.get("caption")
            inputs["language_actions"] = "" if caption is None else decode_text(caption)
            inputs["sample_mask"] = True
            return inputs
        if is_prediction:
            inputs["prompt"] = "predict the robot's action between two images in the prediction"
        fmt = self.language_action_format
        if fmt.include_rotation:
            assert self.action_encoding == 1, "Rotation only supported for EEF_POS encoding"

        if self.transform_strategy == "vla0":  # sample_handlers.py:435-457
            inputs["language_actions"] = fmt.summarize_actions(inputs["actions"]) if "actions" in inputs else ""
            inputs["frame_description"] = "normalized"
            inputs["sample_mask"] = True
            return inputs
        if "language_actions" in data and self.enable_langact_training:  # sample_handlers.py:372-411
            text, frame = self._summarize(data, np.asarray(data["raw_state"]), dataset_name, rotation_applied)
            inputs["language_actions"], inputs["frame_description"] = text, frame
            if self.use_rough_scale:
                inputs["language_actions"] = LA.describe_language_action_scale(text)
                inputs["sample_mask"] = True
            else:
                inputs["sample_mask"] = not LA.is_idle_language_action(text, fmt.get_sum_decimal(), fmt.include_rotation)
            return inputs
        inputs["sample_mask"] = True
        return inputs


@dataclasses.dataclass(frozen=True)
class CoTOutputs:
    """output_transforms.py:20-214: {"actions", ["reasoning", "raw_state"]} -> {"actions", "reasoning"}."""
    language_action_format: LA.LanguageActionFormat | str | None = None
    norm_stats: dict | None = None
    normalization_type: str = "bounds_q99"
    transform_strategy: str = "standard"

    def __post_init__(self):
        if self.language_action_format is not None and not isinstance(self.language_action_format, LA.LanguageActionFormat):
            object.__setattr__(self, "language_action_format", LA.get_language_action_format(self.language_action_format))

    def _unnormalize_vla0(self, actions):
        """:104-183 - [-1, 1] back to physical units over the leading dims the statistics cover."""
        stats = None if self.norm_stats is None else self.norm_stats.get("actions")
        if stats is None:
            return actions
        names, eps = {"bounds_q99": (("q01", "q99"), 1e-6), "bounds": (("min", "max"), 1e-8),
                      "normal": (("mean", "std"), 1e-6)}.get(self.normalization_type, (None, None))
        if names is None:
            return actions
        a, b = (getattr(stats, n, None) for n in names)
        if a is None or b is None:
            return actions
        a, b = np.asarray(a), np.asarray(b)
        d = min(a.shape[-1], actions.shape[-1])
        if self.normalization_type == "normal":
            out = actions[..., :d] * (b[..., :d] + eps) + a[..., :d]
        else:
            out = (actions[..., :d] + 1.0) / 2.0 * (b[..., :d] - a[..., :d] + eps) + a[..., :d]
        return np.concatenate([out, actions[..., d:]], axis=-1) if actions.shape[-1] > d else out

    def __call__(self, data: dict) -> dict:
        if "reasoning" not in data:
            return {"actions": np.asarray(data["actions"]), "reasoning": None}
        reasoning = data.get("reasoning")
        fmt = self.language_action_format
        assert fmt is not None
        assert reasoning is not None
        if self.transform_strategy == "vla0":
            if isinstance(fmt, LA.VLA0ActionFormat):
                return {"actions": self._unnormalize_vla0(fmt.parse_to_full_actions(reasoning)), "reasoning": reasoning}
            movement, gripper = fmt.parse_language_to_deltas(reasoning)
        else:
            state = np.asarray(data["raw_state"]) if fmt.use_eef_frame and "raw_state" in data else None
            movement, gripper = fmt.parse_language_to_deltas(reasoning, initial_state=state)
        return {"actions": movement if gripper is None else np.concatenate([movement, [gripper]]), "reasoning": reasoning}
