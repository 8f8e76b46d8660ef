"""Configuration surface of the hot path, mirroring the reference's dataclasses field-for-field.

Reference:
  - gemma variants ....... src/lap/models/backbones/gemma.py:43-109  (Config, get_config)
  - SigLIP variants ...... third_party/openpi/src/openpi/models/siglip.py:298-373 (decode_variant)
  - LAPConfig ............ src/lap/models/lap_config.py:22-111
  - optimizer / schedule . third_party/openpi/src/openpi/training/optimizer.py:15-109
  - EMA schedule ......... src/lap/training/config.py:372-504,549-589
  - TrainConfig .......... src/lap/training/config.py:507-603, `_CONFIGS` :607-832
Only the fields that reach the hot path are kept; data / checkpoint / wandb fields are out of scope.
"""
from __future__ import annotations

import dataclasses
import math
from typing import Literal

PALIGEMMA_VOCAB_SIZE = 257_152  # gemma.py:40
IMAGE_RESOLUTION = (224, 224)  # OP/models/model.py IMAGE_RESOLUTION


# --------------------------------------------------------------------------------------------
# Gemma (gemma.py:43-109)
# --------------------------------------------------------------------------------------------
@dataclasses.dataclass(frozen=True)
class GemmaConfig:
    width: int
    depth: int
    mlp_dim: int
    num_heads: int
    num_kv_heads: int
    head_dim: int


def get_gemma_config(variant: str) -> GemmaConfig:
    """gemma.py:58-109. LoRA variants are not part of LAP-3B (freeze_filter = Nothing) and are rejected."""
    if variant == "dummy":
        return GemmaConfig(width=64, depth=4, mlp_dim=128, num_heads=8, num_kv_heads=1, head_dim=16)
    if variant == "gemma_300m":
        return GemmaConfig(width=1024, depth=18, mlp_dim=4096, num_heads=8, num_kv_heads=1, head_dim=256)
    if variant == "gemma_2b":
        return GemmaConfig(width=2048, depth=18, mlp_dim=16_384, num_heads=8, num_kv_heads=1, head_dim=256)
    # extra small variants used only by this repo's tests (not in the reference)
    if variant == "tiny_expert":
        return GemmaConfig(width=32, depth=4, mlp_dim=64, num_heads=8, num_kv_heads=1, head_dim=16)
    if variant == "small_2b":  # gemma_2b head geometry (8x256, kv=1) at depth 2 / narrow widths
        return GemmaConfig(width=256, depth=2, mlp_dim=512, num_heads=8, num_kv_heads=1, head_dim=256)
    if variant == "small_300m":
        return GemmaConfig(width=128, depth=2, mlp_dim=256, num_heads=8, num_kv_heads=1, head_dim=256)
    if variant == "mid_2b":  # depth / head geometry of gemma_2b at a quarter of the width (pairs with the real gemma_300m)
        return GemmaConfig(width=512, depth=18, mlp_dim=1024, num_heads=8, num_kv_heads=1, head_dim=256)
    # variants of tests/golden/make_reference_golden.py: the reference's JAX->PyTorch converter only handles
    # width == num_heads * head_dim for the PaliGemma tower (convert_jax_model_to_pytorch.py:205-212)
    if variant == "pin_a":
        return GemmaConfig(width=128, depth=3, mlp_dim=256, num_heads=8, num_kv_heads=1, head_dim=16)
    if variant == "pin_a_expert":
        return GemmaConfig(width=64, depth=3, mlp_dim=128, num_heads=8, num_kv_heads=1, head_dim=16)
    if variant == "pin_b":
        return GemmaConfig(width=256, depth=2, mlp_dim=384, num_heads=8, num_kv_heads=1, head_dim=32)
    if variant == "pin_b_expert":
        return GemmaConfig(width=96, depth=2, mlp_dim=160, num_heads=8, num_kv_heads=1, head_dim=32)
    raise ValueError(f"Unknown variant: {variant}")


# --------------------------------------------------------------------------------------------
# SigLIP (siglip.py:298-373)
# --------------------------------------------------------------------------------------------
@dataclasses.dataclass(frozen=True)
class SiglipConfig:
    width: int
    depth: int
    mlp_dim: int
    num_heads: int
    patch_size: int = 14
    num_classes: int = 2048  # head output = paligemma width (lap.py:78)

    @property
    def head_dim(self) -> int:
        return self.width // self.num_heads


_SIGLIP = {
    "mu": (32, 1, 128, 2),
    "Ti": (192, 12, 768, 3),
    "S": (384, 12, 1536, 6),
    "B": (768, 12, 3072, 12),
    "L": (1024, 24, 4096, 16),
    "So400m": (1152, 27, 4304, 16),
    # test-only: So400m head geometry (hd=72) at depth 2
    "tiny72": (144, 2, 200, 2),
}


def get_siglip_config(variant: str, num_classes: int) -> SiglipConfig:
    v, patch = variant.split("/") if "/" in variant else (variant, "16")
    w, d, m, h = _SIGLIP[v]
    return SiglipConfig(width=w, depth=d, mlp_dim=m, num_heads=h, patch_size=int(patch), num_classes=num_classes)


# --------------------------------------------------------------------------------------------
# LAPConfig (lap_config.py:22-111)
# --------------------------------------------------------------------------------------------
@dataclasses.dataclass(frozen=True)
class LAPConfig:
    dtype: str = "bfloat16"
    paligemma_variant: str = "gemma_2b"
    action_expert_variant: str = "gemma_300m"
    action_dim: int = 7
    action_horizon: int = 16
    max_token_len: int | None = 220
    verbose_mode: bool = False
    pi05: bool = True
    discrete_state_input: bool | None = True
    aug_wrist_image: bool = True
    enable_image_augmentation: bool = True
    use_bimanual: bool = False
    enable_action_training: bool = False
    enable_langact_training: bool = True
    enable_prediction_training: bool = False
    enable_vqa_training: bool = False
    language_loss_weight: float = 1.0
    action_loss_weight: float = 1.0
    prediction_loss_weight: float = 1.0
    vqa_loss_weight: float = 0.1
    stop_action_to_vlm_grad: bool = False
    # host-side (tokenizer / prompt) fields of the reference config, consumed by `make_tokenizer`
    prompt_format: str = "lap"
    prediction_format: str = "default"
    state_dropout: float = 0.0
    reasoning_mask_prob: float = 0.0
    # --- not in the reference: knobs the reference hard-codes, exposed so small test models exist ---
    siglip_variant: str = "So400m/14"  # lap.py:79
    vocab_size: int = PALIGEMMA_VOCAB_SIZE  # lap.py:27
    image_size: int = 224  # OP/models/model.py IMAGE_RESOLUTION

    def __post_init__(self):
        """lap_config.py:76-80: `None` selects the variant-dependent defaults."""
        if self.max_token_len is None:
            object.__setattr__(self, "max_token_len", 200 if self.pi05 else 48)
        if self.discrete_state_input is None:
            object.__setattr__(self, "discrete_state_input", self.pi05)

    def make_tokenizer(self, sp_processor):
        """The tokenizer this config implies (lap_config.py model_transforms: PaligemmaTokenizer(max_token_len, prompt_format,
        prediction_format, reasoning_mask_prob)), around an injected SentencePiece processor."""
        from .tokenizer import CoTTokenizer
        return CoTTokenizer(sp_processor, max_len=self.max_token_len, prompt_format=self.prompt_format,
                            prediction_format=self.prediction_format, reasoning_mask_prob=self.reasoning_mask_prob)

    @property
    def image_keys(self) -> tuple[str, ...]:
        if self.use_bimanual:
            return ("base_0_rgb", "left_wrist_0_rgb", "right_wrist_0_rgb")
        return ("base_0_rgb", "left_wrist_0_rgb")

    @property
    def gemma(self) -> GemmaConfig:
        return get_gemma_config(self.paligemma_variant)

    @property
    def expert(self) -> GemmaConfig:
        return get_gemma_config(self.action_expert_variant)

    @property
    def siglip(self) -> SiglipConfig:
        return get_siglip_config(self.siglip_variant, self.gemma.width)

    @property
    def num_patches(self) -> int:
        return (self.image_size // self.siglip.patch_size) ** 2

    @property
    def prefix_len(self) -> int:
        return len(self.image_keys) * self.num_patches + self.max_token_len

    def create(self, rng=0):
        """lap_config.py:102-111 — returns the engine-backed model (random init from `rng` seed)."""
        from .model import LAP

        return LAP(self, seed=int(rng) if not hasattr(rng, "seed") else rng.seed())

    def load(self, params: dict):
        """OP/models/model.py:233-241 — build the model from a reference-layout parameter tree."""
        from .model import LAP

        model = LAP(self, seed=0, init=False)
        model.load_params(params)
        return model


# --------------------------------------------------------------------------------------------
# Optimizer / schedules (OP/training/optimizer.py)
# --------------------------------------------------------------------------------------------
@dataclasses.dataclass(frozen=True)
class CosineDecaySchedule:
    warmup_steps: int = 1_000
    peak_lr: float = 2.5e-5
    decay_steps: int = 30_000
    decay_lr: float = 2.5e-6

    def lr(self, step: int) -> float:
        """optax.warmup_cosine_decay_schedule(init=peak/(warm+1), peak, warm, decay_steps, end) at `step`.

        optimizer.py:24-31.  optax joins a linear warm-up [0,warm) with cosine_decay_schedule over
        (decay_steps - warm) steps starting at `warm`; alpha = end/peak.
        """
        init = self.peak_lr / (self.warmup_steps + 1)
        if step < self.warmup_steps:
            frac = step / self.warmup_steps if self.warmup_steps > 0 else 1.0
            return init + (self.peak_lr - init) * frac
        decay = max(self.decay_steps - self.warmup_steps, 1)
        count = min(step - self.warmup_steps, decay)
        cosine = 0.5 * (1.0 + math.cos(math.pi * count / decay))
        alpha = self.decay_lr / self.peak_lr if self.peak_lr != 0 else 0.0
        return self.peak_lr * ((1 - alpha) * cosine + alpha)


@dataclasses.dataclass(frozen=True)
class AdamW:
    b1: float = 0.9
    b2: float = 0.95
    eps: float = 1e-8
    weight_decay: float = 1e-10
    clip_gradient_norm: float = 1.0


@dataclasses.dataclass(frozen=True)
class EmaScheduleChoice:
    """src/lap/training/config.py:472-504."""

    kind: Literal["disabled", "constant", "delayed", "cosine_delayed"] = "delayed"
    start_step: int = 10000


@dataclasses.dataclass(frozen=True)
class TrainConfig:
    """Hot-path subset of src/lap/training/config.py:507-603 (same field names)."""

    name: str = "lap"
    model: LAPConfig = dataclasses.field(default_factory=LAPConfig)
    lr_schedule: CosineDecaySchedule = dataclasses.field(  # build_cosine_lr() defaults, config.py:41-55
        default_factory=lambda: CosineDecaySchedule(warmup_steps=5000, peak_lr=1e-4, decay_steps=40_000, decay_lr=1e-4)
    )
    optimizer: AdamW = dataclasses.field(default_factory=lambda: AdamW(weight_decay=0.0001))
    num_train_steps: int = 40_000
    batch_size: int = 32  # OP/training/config.py:497
    log_interval: int = 50
    save_interval: int = 1000
    keep_period: int | None = 5000
    seed: int = 0
    fsdp_devices: int = 1
    ema_decay: float | None = 0.999
    ema_schedule_choice: EmaScheduleChoice = dataclasses.field(
        default_factory=lambda: EmaScheduleChoice(kind="cosine_delayed", start_step=5000)
    )

    def get_ema_init(self) -> tuple[float | None, bool]:
        """config.py:554-563."""
        k = self.ema_schedule_choice.kind
        if k == "cosine_delayed":
            return (None, False) if self.ema_decay is None else (0.0, True)
        if k == "disabled" or self.ema_decay is None:
            # schedule is None -> (ema_decay, ema_decay is not None)
            return self.ema_decay, self.ema_decay is not None
        if k == "constant":
            return self.ema_decay, True
        if k == "delayed":
            if self.ema_schedule_choice.start_step <= 0:
                return self.ema_decay, True
            return None, True
        raise ValueError(k)

    def get_ema_decay_for_step(self, step: int) -> tuple[float, bool]:
        """config.py:565-589 -> (decay, enabled)."""
        k = self.ema_schedule_choice.kind
        if k == "cosine_delayed":
            if self.ema_decay is None:
                return 0.0, False
            start = self.ema_schedule_choice.start_step
            duration = max(self.num_train_steps - start, 1)
            progress = min(max((step - start) / duration, 0.0), 1.0)
            return self.ema_decay * (1.0 - math.cos(math.pi * progress)) / 2.0, step >= start
        if self.ema_decay is None:
            return 0.0, False
        if k == "disabled":
            # schedule None and ema_decay not None -> constant decay, enabled (config.py:587-589)
            return float(self.ema_decay), True
        if k == "constant":
            return float(self.ema_decay), True
        if k == "delayed":
            start = self.ema_schedule_choice.start_step
            if step >= start:
                return float(self.ema_decay), True
            return 0.0, False
        raise ValueError(k)


_LR = CosineDecaySchedule(warmup_steps=1000, peak_lr=5e-5, decay_steps=40_000, decay_lr=5e-5)

_CONFIGS = {
    # src/lap/training/config.py:608-619
    "lap": TrainConfig(
        name="lap",
        model=LAPConfig(action_dim=7, action_horizon=16, max_token_len=180, enable_action_training=True,
                        stop_action_to_vlm_grad=True),
        batch_size=2048,  # everything else is the TrainConfig default (cosine LR 1e-4 / 5000 warm-up, cosine-delayed EMA from 5000)
    ),
    # src/lap/training/config.py:751-785
    "lap_libero": TrainConfig(
        name="lap_libero",
        model=LAPConfig(action_dim=7, action_horizon=10, max_token_len=180, enable_action_training=True,
                        stop_action_to_vlm_grad=False, language_loss_weight=0.4, enable_image_augmentation=False),
        lr_schedule=_LR, num_train_steps=40_001, batch_size=256, save_interval=2000, keep_period=2000,
        ema_schedule_choice=EmaScheduleChoice(kind="constant"),
    ),
    # test-only: tiny models that exercise the same code paths
    "debug_tiny": TrainConfig(
        name="debug_tiny",
        model=LAPConfig(paligemma_variant="dummy", action_expert_variant="tiny_expert", siglip_variant="mu/14",
                        action_dim=7, action_horizon=10, max_token_len=24, enable_action_training=True,
                        language_loss_weight=0.4, enable_image_augmentation=False, vocab_size=512, image_size=56),
        lr_schedule=CosineDecaySchedule(warmup_steps=2, peak_lr=1e-3, decay_steps=10, decay_lr=1e-4),
        num_train_steps=10, batch_size=4, ema_schedule_choice=EmaScheduleChoice(kind="constant"),
    ),
    "debug_small": TrainConfig(
        name="debug_small",
        model=LAPConfig(paligemma_variant="small_2b", action_expert_variant="small_300m", siglip_variant="tiny72/14",
                        action_dim=7, action_horizon=10, max_token_len=40, enable_action_training=True,
                        language_loss_weight=0.4, enable_image_augmentation=False, vocab_size=2048, image_size=112),
        lr_schedule=CosineDecaySchedule(warmup_steps=2, peak_lr=1e-3, decay_steps=10, decay_lr=1e-4),
        num_train_steps=10, batch_size=4, ema_schedule_choice=EmaScheduleChoice(kind="constant"),
    ),
    # BASELINE.json's 48-token / 50-step variant (upstream Pi0Config defaults, OP/models/pi0_config.py:25-37) at test size
    "debug_bj": TrainConfig(
        name="debug_bj",
        model=LAPConfig(paligemma_variant="small_2b", action_expert_variant="small_300m", siglip_variant="tiny72/14",
                        action_dim=32, action_horizon=50, max_token_len=48, enable_action_training=True,
                        language_loss_weight=1.0, enable_image_augmentation=False, vocab_size=1024, image_size=56),
        lr_schedule=CosineDecaySchedule(warmup_steps=2, peak_lr=1e-3, decay_steps=10, decay_lr=1e-4),
        num_train_steps=10, batch_size=4, ema_schedule_choice=EmaScheduleChoice(kind="cosine_delayed", start_step=2),
    ),
}


def get_config(name: str) -> TrainConfig:
    """src/lap/training/config.py:843-862."""
    if name not in _CONFIGS:
        raise ValueError(f"Config '{name}' not found. Available: {sorted(_CONFIGS)}")
    return _CONFIGS[name]
