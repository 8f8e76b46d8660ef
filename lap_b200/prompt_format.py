"""Prompt strings of the lang-action sample (third slice): what `PromptFormat.format_prompt`
(src/lap/models/prompt_utils/prompt.py:115-164) produces for the formats the training / serving configs select —
"lap" (:167-179), "vla0_chunked" (:216-230), "default_prediction" (:182-199), "default_vqa" (:202-210) — including the state
discretisation of `StateDiscretizationConfig.discretize_state` (prompt_utils/state.py:125-160: trim trailing zero padding but
keep >= 10 dims, 256 left-closed bins on [-1, 1), `np.digitize - 1`) and the task clean-up of `TaskModule.format_task`
(prompt.py:44-61).

The reference composes four module objects per format; here a format is ONE flat record (the modules only ever carry strings),
and the output strings are compared byte for byte with the reference classes executed from source
(tests/golden/make_reference_prompt_golden.py -> tests/golden/reference_prompts.json.gz).
"""
from __future__ import annotations

import dataclasses
import random
import re
from collections.abc import Callable

import numpy as np

_DIRECTION_WORDS = ("right", "left", "forward", "up", "down", "back", "clockwise", "counterclockwise")
_STATE_TYPE_LABELS = {"joint_pos": " (joint position)", "eef_pose": " (end-effector pose)"}


# ---- token-piece predicates (prompt_utils/checkers.py) ----
def is_number(piece: str) -> bool:
    return re.search(r"[0-9]", piece) is not None


def is_direction_natural(piece: str) -> bool:
    low = piece.lower()
    return any(w in low for w in _DIRECTION_WORDS)


def is_direction_schema(piece: str) -> bool:
    return "+" in piece or "-" in piece


def is_direction_none(piece: str) -> bool:
    return False


def is_critical_directional(piece: str) -> bool:
    return is_number(piece) or is_direction_natural(piece)


def is_critical_schema(piece: str) -> bool:
    return is_number(piece) or is_direction_schema(piece)


def discretize_state(state, bins: int = 256, min_dim: int = 10, range_min: float = -1.0, range_max: float = 1.0) -> str:
    """state.py:125-160 with the default (space separated) template.  Values below `range_min` land in bin -1 and values
    >= the last edge in bin `bins - 1`, exactly as `np.digitize(x, edges) - 1` does."""
    x = np.asarray(state)
    live = np.abs(x.reshape(-1, x.shape[-1]) if x.ndim > 1 else x[None]) > 1e-8
    cols = np.flatnonzero(live.any(axis=0))
    keep = max(int(cols[-1]) + 1 if cols.size else 0, min_dim)
    x = x[..., :keep].reshape(-1)
    if x.size == 0:
        return ""
    edges = np.linspace(range_min, range_max, bins + 1)[:-1]
    return " ".join(str(v) for v in np.digitize(x, bins=edges) - 1)


@dataclasses.dataclass(frozen=True)
class PromptFormat:
    name: str
    prefix: str | None = None                     # PrefixModule.text
    task_template: str | None = "Task: {prompt}, predict the robot's action in the {frame_description}"
    include_time_horizon: bool = False
    time_horizon_template: str = ("predict the robot's action in the future {time_horizon_seconds} seconds in the "
                                  "{frame_description}")
    state_template: str | None = None             # StateModule.state_prefix_template; None = format carries no state
    state_bins: int = 256
    include_state_type: bool = False
    action_prefix: str | None = "Answer: "
    separator: str = ""
    critical_token_checker: Callable[[str], bool] | None = is_number
    direction_token_checker: Callable[[str], bool] | None = is_direction_none

    @property
    def include_state(self) -> bool:
        return self.state_template is not None

    def _task(self, prompt, time_horizon_seconds, frame_description) -> str:
        text = prompt.strip().replace("_", " ").replace("\n", " ").rstrip(".")
        if self.include_time_horizon:
            assert time_horizon_seconds is not None, "Time horizon must be provided if include_time_horizon is True"
            # NB the default horizon template also names {frame_description}, which the reference does not pass here
            # (prompt.py:59): like there, that template raises KeyError; a custom template without it works
            text += ", " + self.time_horizon_template.format(time_horizon_seconds=round(time_horizon_seconds * 2) / 2.0)
        return self.task_template.format(prompt=text, frame_description=frame_description)

    def _state(self, state, state_type) -> str:
        if state is None or state_type == "none":
            return self.state_template.format(state="", state_label="None" if self.include_state_type else "")
        label = (_STATE_TYPE_LABELS.get(state_type, state_type) if state_type else "") if self.include_state_type else ""
        return self.state_template.format(state=discretize_state(state, bins=self.state_bins), state_label=label)

    def format_prompt(self, prompt: str, state=None, state_type: str | None = None,
                      time_horizon_seconds: float | None = None, frame_description: str = "robot base frame",
                      state_dropout: float = 0.0) -> str:
        parts = []
        if self.prefix is not None:
            parts.append(self.prefix)
        if self.task_template is not None:
            parts.append(self._task(prompt, time_horizon_seconds, frame_description))
        # same short-circuit order as prompt.py:147-149: `random.random()` is drawn only when a state would be added
        if not (self.state_template is None or state is None or (state_dropout > 0.0 and random.random() < state_dropout)):
            s = self._state(state, state_type)
            if s:
                parts.append(s)
        if self.action_prefix is not None:
            parts.append(self.action_prefix)
        return self.separator.join(parts)


_STATE = "State{state_label}: {state}"
LAP_PROMPT_FORMAT = PromptFormat(name="lap", state_template=_STATE, separator="; ",
                                 critical_token_checker=is_critical_directional, direction_token_checker=is_direction_natural)
DEFAULT_PREDICTION_PROMPT_FORMAT = PromptFormat(name="default_prediction", task_template="Task: {prompt}",
                                                state_template=_STATE, separator="; ",
                                                critical_token_checker=is_critical_schema,
                                                direction_token_checker=is_direction_schema)
DEFAULT_VQA_PROMPT_FORMAT = PromptFormat(name="default_vqa", task_template="Task: {prompt}", separator="; ",
                                         critical_token_checker=None, direction_token_checker=None)
VLA0_CHUNKED_PROMPT_FORMAT = PromptFormat(
    name="vla0_chunked",
    prefix=("Analyze the input image and predict robot actions for the next 10 timesteps. "
            "Each action has 7 dimensions. Output a single sequence of 70 integers (0-1000 each), "
            "representing the 10 timesteps sequentially. Provide only space-separated numbers. Nothing else."),
    task_template="Task: {prompt}", action_prefix="", separator="\n", critical_token_checker=is_number,
    direction_token_checker=is_direction_none)

PROMPT_FORMAT_REGISTRY = {"lap": LAP_PROMPT_FORMAT, "vla0_chunked": VLA0_CHUNKED_PROMPT_FORMAT}


def _question_format(name, checker=None):
    """prompt.py:236-315: the six auxiliary question types share one layout and differ in their checkers only."""
    return PromptFormat(name=name, task_template="Task: {prompt}", separator="; ", critical_token_checker=checker,
                        direction_token_checker=checker)


PREDICTION_PROMPT_FORMAT_REGISTRY = {
    "default": DEFAULT_PREDICTION_PROMPT_FORMAT,
    "task_prediction": _question_format("task_prediction"),
    "direction_classification": _question_format("direction_classification", is_direction_natural),
    "gripper_prediction": _question_format("gripper_prediction"),
    "magnitude_estimation": _question_format("magnitude_estimation"),
    "temporal_ordering": _question_format("temporal_ordering"),
    "embodiment_identification": _question_format("embodiment_identification"),
}
VQA_PROMPT_FORMAT_REGISTRY = {"default_vqa": DEFAULT_VQA_PROMPT_FORMAT}
