"""Language actions (fourth slice): numeric end-effector deltas <-> the text the lang-action head is trained on.

Reference: src/lap/policies/transforms/action_text.py (`summarize_numeric_actions` :46-140, compact :25-43, bimanual :186-207,
`describe_language_action_scale` :143-183, `is_idle_language_action` :210-302), frame_transforms.py (base <-> end-effector
frame, :22-129), lang_action_formats.py (`LanguageActionFormat.parse_language_to_deltas` :38-131, `VLA0ActionFormat` :134-263,
registry :266-307).  Host-side numpy / regex work; scipy's `Rotation` does the Euler conversions exactly as in the reference.
Strings compare byte for byte and arrays bit for bit with the reference functions executed from source
(tests/golden/make_reference_langaction_golden.py -> tests/golden/reference_langactions.json.gz).
"""
from __future__ import annotations

import dataclasses
import logging
import re
from typing import Literal

import numpy as np

_DEG = 180.0 / np.pi
_MOVE_AXIS = {"forward": (0, 1.0), "backward": (0, -1.0), "back": (0, -1.0), "left": (1, 1.0), "right": (1, -1.0),
              "up": (2, 1.0), "down": (2, -1.0)}
_ROT_WORDS = r"(tilt left|tilt right|tilt up|tilt down|tilt back|tilt forward|rotate clockwise|rotate counterclockwise)"


# ------------------------------------------------------------------ frames (frame_transforms.py)
def rot6d_to_rotmat(rot6d) -> np.ndarray:
    """:7-19 - Gram-Schmidt on the two 3-vectors; the basis vectors are the COLUMNS of the result."""
    r = np.asarray(rot6d)
    a1, a2 = r[..., 0:3], r[..., 3:6]
    b1 = a1 / np.linalg.norm(a1, axis=-1, keepdims=True)
    a2 = a2 - np.sum(b1 * a2, axis=-1, keepdims=True) * b1
    b2 = a2 / np.linalg.norm(a2, axis=-1, keepdims=True)
    return np.stack([b1, b2, np.cross(b1, b2, axis=-1)], axis=-1)


def _R():
    from scipy.spatial.transform import Rotation
    return Rotation


def transform_actions_to_eef_frame(actions, initial_state, dataset_name, needs_wrist_rotation: bool = False) -> np.ndarray:
    """:22-70 - one [>=6] delta from the robot base frame into the end-effector frame of `initial_state`
    (xyz, rot6d at [3:9]), with the per-dataset axis conventions."""
    a = np.asarray(actions, dtype=float)
    s = np.asarray(initial_state, dtype=float)
    assert a.ndim == 1
    out = a.copy()
    to_eef = rot6d_to_rotmat(s[3:9]).T
    p = to_eef @ a[:3]
    p[1], p[2] = -p[1], -p[2]
    if "jaco_play" in dataset_name:
        p = np.array([p[1], p[0], -p[2]])
    elif "berkeley_autolab_ur5" in dataset_name:
        p = np.array([-p[1], p[0], p[2]])
    out[:3] = p
    R = _R()
    e = R.from_matrix(to_eef @ R.from_euler("xyz", a[3:6]).as_matrix() @ to_eef.T).as_euler("xyz")
    if not needs_wrist_rotation:
        e[1], e[2] = -e[1], -e[2]
    if any(k in dataset_name for k in ("furniture_bench_dataset_converted_externally_to_rlds", "austin", "fmb", "viola")):
        e[1], e[2] = -e[1], -e[2]
    elif "berkeley_autolab_ur5" in dataset_name:
        e[1] = -e[1]
    out[3:6] = e
    return out


def transform_actions_from_eef_frame(actions, initial_state, dataset_name: str = "") -> np.ndarray:
    """:73-129 - the way back, for [T, >=3] deltas; `initial_state` is xyz + Euler (len 7) or xyz + rot6d."""
    a = np.asarray(actions, dtype=float)
    s = np.asarray(initial_state, dtype=float)
    if s.ndim == 2:
        assert s.shape[0] == 1
        s = s[0]
    if a.ndim == 1:
        a = a[None, :]
    out = a.copy()
    R = _R()
    to_base = R.from_euler("xyz", s[3:6]).as_matrix() if len(s) == 7 else rot6d_to_rotmat(s[3:9])
    for i in range(len(out)):
        p = a[i, :3].copy()
        if "jaco_play" in dataset_name:
            p = np.array([p[1], p[0], -p[2]])
        elif "berkeley_autolab" in dataset_name:
            p = np.array([p[1], -p[0], p[2]])
        else:
            p[1], p[2] = -p[1], -p[2]
        out[i, :3] = to_base @ p
        if a.shape[-1] >= 6:
            e = a[i, 3:6].copy()
            if "furniture_bench" in dataset_name or "utaustin" in dataset_name or "fmb" in dataset_name:
                e[1], e[2] = -e[1], -e[2]
            elif "berkeley_autolab" in dataset_name:
                e[1] = -e[1]
            elif "jaco_play" not in dataset_name:
                e[1], e[2] = -e[1], -e[2]
            out[i, 3:6] = R.from_matrix(to_base @ R.from_euler("xyz", e).as_matrix() @ to_base.T).as_euler("xyz")
    return out


# ------------------------------------------------------------------ deltas -> text (action_text.py)
def _nearest(value: float, n: int = 5) -> int:
    return int(round(value / n) * n)


def _decimals(sum_decimal: str) -> int:
    m = re.fullmatch(r"(\d+)f", sum_decimal)
    return int(m.group(1)) if m else 0


def _summed(arr_like):
    arr = np.asarray(arr_like, dtype=float)
    return arr[None, :] if arr.ndim == 1 else arr


def _compact(arr, include_rotation: bool) -> str:
    """:25-43 - "<+03 -01 +00 [+05 +00 -10] 1>": summed cm, degrees to the nearest 5, last gripper bit."""
    parts = [f"{int(round(float(arr[..., k].sum()) * 100.0)):+03d}" for k in range(3)]
    if include_rotation:
        parts += [f"{_nearest(float(arr[..., k].sum()) * 180.0 / np.pi, 5):+03d}" for k in (3, 4, 5)]
    parts.append(str(1 if float(arr[-1, 6]) >= 0.5 else 0))
    return "<" + " ".join(parts) + ">"


def summarize_numeric_actions(arr_like, sum_decimal: str, include_rotation: bool = False, rotation_precision: int = 10):
    """:46-140 - chunk of [T, >=7] deltas (m, rad, gripper) -> "move forward 3 cm, move up 1 cm, ..., open gripper".
    Order of the clauses: x, z, y for the numeric styles; x, y, z for "no_number" (as in the reference)."""
    arr = _summed(arr_like)
    if arr.shape[-1] < 7:
        return None
    if sum_decimal == "compact":
        return _compact(arr, include_rotation)
    numbered = sum_decimal != "no_number"
    dec = _decimals(sum_decimal)
    d_m = [float(arr[..., k].sum()) for k in range(3)]
    mag = [round(abs(v * 100.0), dec) for v in d_m]
    r_rad = [float(arr[..., k].sum()) for k in (3, 4, 5)] if include_rotation else []
    r_mag = [_nearest(abs(v * 180.0 / np.pi), rotation_precision) for v in r_rad]

    def num(v):
        if sum_decimal == "nearest_10":
            return str(int(round(v / 10) * 10))
        return f"{v:.{dec}f}"

    parts = []
    names = (("forward", "back"), ("left", "right"), ("up", "down"))
    for k in ((0, 2, 1) if numbered else (0, 1, 2)):
        if mag[k] != 0 and d_m[k] != 0:
            word = names[k][0] if d_m[k] > 0 else names[k][1]
            parts.append(f"move {word} {num(mag[k])} cm" if numbered else f"move {word}")
    rot_names = (("tilt left", "tilt right"), ("tilt back", "tilt forward"), ("rotate counterclockwise", "rotate clockwise"))
    for k, v in enumerate(r_rad):
        if v != 0 and (r_mag[k] != 0 or not numbered):
            word = rot_names[k][0] if v > 0 else rot_names[k][1]
            parts.append(f"{word} {r_mag[k]} degrees" if numbered else word)
    parts.append("open gripper" if float(arr[-1, 6]) >= 0.5 else "close gripper")
    return ", ".join(parts)


def summarize_bimanual_numeric_actions(arr_like, sum_decimal: str, include_rotation: bool = False):
    """:186-207 - two 7-dim arms side by side."""
    arr = _summed(arr_like)
    if arr.shape[-1] < 14:
        return None
    left, right = arr[..., :7], arr[..., 7:14]
    if sum_decimal == "compact":
        return f"<L {_compact(left, include_rotation)[1:-1]} R {_compact(right, include_rotation)[1:-1]}>"
    ls = summarize_numeric_actions(left, sum_decimal, include_rotation)
    rs = summarize_numeric_actions(right, sum_decimal, include_rotation)
    return None if ls is None or rs is None else f"Left arm: {ls}. Right arm: {rs}"


_SCALE_T = re.compile(r"(move\s+(?:forward|back|left|right|up|down))\s+([+\-]?\d+(?:\.\d+)?)\s*cm")
_SCALE_R = re.compile(r"((?:tilt\s+(?:left|right|back|forward))|(?:rotate\s+(?:clockwise|counterclockwise)))\s+"
                      r"([+\-]?\d+(?:\.\d+)?)\s*degrees")


def describe_language_action_scale(language_action):
    """:143-183 - numbers -> "slightly" / "moderately" / "a lot" (<=3 / <8 cm; <10 / <30 degrees)."""
    if language_action is None:
        return None
    if not isinstance(language_action, str) or not language_action.strip():
        return language_action

    def sub(pattern, text, lo, mid, lo_inclusive):
        def rep(m):
            v = float(m.group(2))
            word = "slightly" if (v <= lo if lo_inclusive else v < lo) else ("moderately" if v < mid else "a lot")
            return f"{m.group(1)} {word}"
        return pattern.sub(rep, text)

    return sub(_SCALE_R, sub(_SCALE_T, language_action, 3.0, 8.0, True), 10.0, 30.0, False)


def _parse_moves(text, pattern):
    d = [0.0, 0.0, 0.0]
    for m in pattern.finditer(text):
        axis, sign = _MOVE_AXIS[m.group(1).lower()]
        d[axis] += sign * (float(m.group(2)) if m.group(2) is not None else 0.0)
    return d


def _parse_rotations(text, pitch_up_sign):
    """Sums "tilt/rotate ... N degrees" clauses into (roll, pitch, yaw) degrees.  `pitch_up_sign` carries a reference quirk:
    `parse_language_to_deltas` counts "tilt up/forward" as NEGATIVE pitch, `is_idle_language_action` as positive."""
    r = [0.0, 0.0, 0.0]
    for m in re.finditer(_ROT_WORDS + r"\s+([\d.]+)\s*degrees", text, re.IGNORECASE):
        kind, v = m.group(1).lower(), float(m.group(2))
        if kind in ("tilt left", "tilt right"):
            r[0] += v if kind == "tilt left" else -v
        elif kind in ("tilt up", "tilt forward"):
            r[1] += pitch_up_sign * v
        elif kind in ("tilt down", "tilt back"):
            r[1] -= pitch_up_sign * v
        else:
            r[2] += v if kind == "rotate counterclockwise" else -v
    return r


def is_idle_language_action(language_action, sum_decimal: str, include_rotation: bool = False,
                            translation_threshold: float = 1.0, rotation_threshold_deg: float = 10.0) -> bool:
    """:210-302 - True when the text moves < 1 cm (and turns < 10 degrees): such samples get `sample_mask` False."""
    if not language_action or not isinstance(language_action, str):
        return True
    if sum_decimal == "compact":
        n = 6 if include_rotation else 3
        m = re.search("<" + r"\s+".join([r"([+\-]\d+)"] * n) + r"\s+\d>", language_action)
        if not m:
            return True
        v = [int(g) for g in m.groups()]
        still = np.sqrt(v[0] ** 2 + v[1] ** 2 + v[2] ** 2) < translation_threshold
        return bool(still and np.sqrt(v[3] ** 2 + v[4] ** 2 + v[5] ** 2) < rotation_threshold_deg) if include_rotation \
            else bool(still)
    if sum_decimal == "no_number":
        moved = re.search(r"move\s+(right|left|forward|backward|back|up|down)(?!\s+[\d.])", language_action, re.IGNORECASE)
        turned = include_rotation and re.search(_ROT_WORDS + r"(?!\s+[\d.])", language_action, re.IGNORECASE)
        return not (moved or turned)
    d = _parse_moves(language_action, re.compile(r"move\s+(right|left|forward|backward|back|up|down)\s+([\d.]+)\s*cm", re.IGNORECASE))
    still = np.sqrt(d[0] ** 2 + d[1] ** 2 + d[2] ** 2) < translation_threshold
    if not include_rotation:
        return bool(still)
    r = _parse_rotations(language_action, +1.0)
    return bool(still and np.sqrt(r[0] ** 2 + r[1] ** 2 + r[2] ** 2) < rotation_threshold_deg)


# ------------------------------------------------------------------ formats (lang_action_formats.py)
@dataclasses.dataclass(frozen=True)
class LanguageActionFormat:
    name: str
    style: Literal["verbose", "compact", "vla0"] = "verbose"
    decimal_places: int = 0
    include_rotation: bool = False
    translation_unit: str = "cm"
    use_eef_frame: bool = False

    def get_sum_decimal(self) -> str:
        return "compact" if self.style == "compact" else f"{self.decimal_places}f"

    def parse_language_to_deltas(self, reasoning, *, initial_state=None):
        """:38-131 - text -> ([dx, dy, dz, droll, dpitch, dyaw] in m / rad, gripper | None); back to the base frame when the
        format is end-effector relative and a state is given.  (The compact style only parses its 7-field rotation form.)"""
        movement = np.zeros(6, dtype=float)
        gripper = None
        if self.style == "compact":
            if self.include_rotation:
                m = re.search(r"<" + r"\s+".join([r"([+\-]\d+)"] * 6) + r"\s+(\d)>", reasoning)
                if m:
                    g = m.groups()
                    movement[:3] = np.array(g[0:3], dtype=float) / 100.0
                    movement[3:6] = np.array(g[3:6], dtype=float) * np.pi / 180.0
                    gripper = float(g[-1])
        else:
            reasoning = reasoning.replace("slightly", "1.5 cm").replace("moderately", "5 cm").replace("a lot", "10 cm")
            d = _parse_moves(reasoning, re.compile(
                rf"move\s+(right|left|forward|backward|back|up|down)(?:\s+([\-\d\.]+)\s*{self.translation_unit})?", re.IGNORECASE))
            movement[:3] = np.array(d, dtype=float) / 100.0
            if self.include_rotation:
                r = _parse_rotations(reasoning, -1.0)
                movement[3:6] = [r[0] * np.pi / 180.0, r[1] * np.pi / 180.0, r[2] * np.pi / 180.0]
            low = reasoning.lower()
            gm = re.search(r"set\s+gripper\s+to\s+([\-+]?\d+\.?\d*)", reasoning, re.IGNORECASE)
            if "open gripper" in low:
                gripper = 1.0
            elif "close gripper" in low:
                gripper = 0.0
            elif gm:
                gripper = float(gm.group(1))
        if self.use_eef_frame and initial_state is not None:
            movement = transform_actions_from_eef_frame(movement, initial_state)[0]
        return movement, gripper


@dataclasses.dataclass(frozen=True)
class VLA0ActionFormat(LanguageActionFormat):
    """:134-263 - VLA-0: normalised actions as space-separated integers in [0, num_bins]."""
    name: str = "vla0"
    style: Literal["vla0"] = "vla0"
    num_bins: int = 1000
    action_horizon: int = 1
    action_dim: int = 7

    def get_sum_decimal(self) -> str:
        return "vla0"

    def summarize_actions(self, actions) -> str:
        a = np.clip(np.atleast_2d(np.asarray(actions, dtype=float)), -1.0, 1.0)
        q = np.clip(np.round((a + 1.0) / 2.0 * self.num_bins).astype(int), 0, self.num_bins)
        return " ".join(map(str, q.flatten()))

    def _grid(self, ints):
        x = np.array(ints, dtype=float) / self.num_bins * 2.0 - 1.0
        n = self.action_horizon * self.action_dim
        x = np.pad(x, (0, n - len(x))) if len(x) < n else x[:n]
        return x.reshape(self.action_horizon, self.action_dim)

    def parse_language_to_deltas(self, reasoning, *, initial_state=None):
        if isinstance(reasoning, list):
            reasoning = " ".join(reasoning)
        try:
            ints = [int(x) for x in reasoning.split()]
        except ValueError:
            return np.zeros(6, dtype=float), None
        if not ints:
            return np.zeros(6, dtype=float), None
        a = self._grid(ints)
        return (a[0, :6] if a.shape[1] >= 6 else np.zeros(6)), (float(a[0, 6]) if a.shape[1] >= 7 else None)

    def parse_to_full_actions(self, reasoning) -> np.ndarray:
        if isinstance(reasoning, list):
            reasoning = " ".join(reasoning)
        zeros = np.zeros((self.action_horizon, self.action_dim), dtype=float)
        if not re.search(r"([\d\s]+)", reasoning):
            logging.info(f"No match found for VLA0 format: {reasoning}")
            return zeros
        try:
            ints = [int(x) for x in reasoning.split()]
        except ValueError:
            logging.info(f"Failed to parse VLA0 format: {reasoning}")
            return zeros
        return self._grid(ints) if ints else zeros


VERBOSE_WITH_ROTATION_FORMAT = LanguageActionFormat(name="verbose_with_rotation", include_rotation=True)
VERBOSE_EEF_WITH_ROTATION_FORMAT = LanguageActionFormat(name="verbose_eef_with_rotation", include_rotation=True, use_eef_frame=True)
VLA0_CHUNKED_FORMAT = VLA0ActionFormat(name="vla0_chunked", num_bins=1000, action_horizon=10, action_dim=7)
LANGUAGE_ACTION_FORMAT_REGISTRY = {f.name: f for f in (VERBOSE_WITH_ROTATION_FORMAT, VERBOSE_EEF_WITH_ROTATION_FORMAT,
                                                       VLA0_CHUNKED_FORMAT)}


def get_language_action_format(name: str) -> LanguageActionFormat:
    if name not in LANGUAGE_ACTION_FORMAT_REGISTRY:
        raise ValueError(f"Unknown language action format: {name}. Available formats: "
                         f"{list(LANGUAGE_ACTION_FORMAT_REGISTRY.keys())}")
    return LANGUAGE_ACTION_FORMAT_REGISTRY[name]
