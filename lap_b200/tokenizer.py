"""Token / mask producer of the lang-action prompt (second slice): the index arithmetic of
`PaligemmaTokenizer.tokenize` (src/lap/models/tokenizer.py:237-315) and its helpers `_create_base_masks` (:105-138),
`_apply_reasoning_dropout` (:140-172), `_build_number_direction_masks` (:174-207), `is_number`
(src/lap/models/prompt_utils/checkers.py:4-6).

What is injected instead of restated: the SentencePiece processor (the reference downloads
gs://big_vision/paligemma_tokenizer.model; any `sentencepiece.SentencePieceProcessor` works).  The prompt format is a
registry name or an object with `format_prompt(...) -> str` and `direction_token_checker(piece) -> bool`
(`lap_b200.prompt_format`, or the reference's own `PromptFormat`).  Everything downstream of the formatted string — BOS/EOS
placement, truncation, `tokenized_prompt_mask`, `tokenized_langact_mask`, `token_loss_mask`, number / direction masks,
right padding — is bit-exact against the reference method executed from source
(tests/golden/make_reference_tokenizer_golden.py -> tests/golden/reference_tokenizer.npz).
"""
from __future__ import annotations

import numpy as np

from .prompt_format import (DEFAULT_VQA_PROMPT_FORMAT, PREDICTION_PROMPT_FORMAT_REGISTRY, PROMPT_FORMAT_REGISTRY,
                            is_direction_natural, is_number)  # noqa: F401  (re-exported)


def _resolve(fmt, registry, what):
    """tokenizer.py:51-71: a registry name or a format object."""
    if isinstance(fmt, str):
        if fmt not in registry:
            raise ValueError(f"Unknown {what}: {fmt}. Available formats: {list(registry.keys())}")
        return registry[fmt]
    return fmt


class CoTTokenizer:
    def __init__(self, sp_processor, max_len: int = 48, prompt_format="lap", prediction_format="default",
                 reasoning_mask_prob: float = 0.0):
        """tokenizer.py:221-235 with the processor passed in (`sentencepiece.SentencePieceProcessor(model_file=...)`)."""
        self._tokenizer = sp_processor
        self._max_len = int(max_len)
        self._prompt_format = _resolve(prompt_format, PROMPT_FORMAT_REGISTRY, "prompt format")
        self._prediction_format = _resolve(prediction_format, PREDICTION_PROMPT_FORMAT_REGISTRY, "prediction format")
        self._vqa_format = DEFAULT_VQA_PROMPT_FORMAT
        self.reasoning_mask_prob = float(reasoning_mask_prob)

    # tokenizer.py:93-103
    def _resolve_format(self, is_vqa_sample: bool, is_prediction_sample: bool):
        if is_prediction_sample:
            return self._prediction_format
        if is_vqa_sample:
            return self._vqa_format
        return self._prompt_format

    # tokenizer.py:105-138
    def _create_base_masks(self, token_count: int, reasoning_start: int, reasoning_end: int, has_reasoning: bool):
        attn_mask = np.zeros(self._max_len, dtype=bool)
        token_loss_mask = np.ones(self._max_len, dtype=bool)
        attn_mask[:token_count] = True
        if not has_reasoning:
            return attn_mask, None, token_loss_mask
        reasoning_mask = np.zeros(self._max_len, dtype=bool)
        start_idx = max(0, min(self._max_len, reasoning_start))
        end_idx = max(0, min(self._max_len, reasoning_end))
        if end_idx > start_idx:
            reasoning_mask[start_idx:end_idx] = True
        return attn_mask, reasoning_mask, token_loss_mask

    # tokenizer.py:140-172 (draws from numpy's global RNG exactly like the reference)
    def _apply_reasoning_dropout(self, token_loss_mask, reasoning_mask, is_vqa_sample: bool):
        if not 0.0 <= self.reasoning_mask_prob <= 1.0:
            raise ValueError(f"reasoning_mask_prob must be between 0.0 and 1.0, got {self.reasoning_mask_prob}")
        if self.reasoning_mask_prob <= 0.0 or is_vqa_sample:
            return token_loss_mask
        reasoning_indices = np.where(reasoning_mask)[0]
        if len(reasoning_indices) == 0:
            return token_loss_mask
        drop_mask = np.random.rand(len(reasoning_indices)) < self.reasoning_mask_prob
        if np.any(drop_mask):
            token_loss_mask[reasoning_indices[drop_mask]] = False
        return token_loss_mask

    # tokenizer.py:174-207
    def _build_number_direction_masks(self, tokens, reasoning_mask, fmt, is_vqa_sample: bool):
        number_mask = np.zeros(self._max_len, dtype=bool)
        direction_mask = np.zeros(self._max_len, dtype=bool)
        if is_vqa_sample:
            return number_mask, direction_mask
        for i in np.where(reasoning_mask)[0]:
            piece = self._tokenizer.id_to_piece(int(tokens[i]))
            if piece:
                if is_number(piece):
                    number_mask[i] = True
                if fmt.direction_token_checker(piece):
                    direction_mask[i] = True
        return number_mask, direction_mask

    # tokenizer.py:237-315
    def tokenize(self, prompt: str, reasoning: str | None = None, state=None, state_type: str | None = None, *,
                 is_vqa_sample: bool = False, is_prediction_sample: bool = False,
                 time_horizon_seconds: float | None = None, frame_description: str = "robot base frame",
                 state_dropout: float = 0.0):
        """-> (tokens int32[max_len], attn_mask, reasoning_mask | None, number_mask | None, direction_mask | None,
        token_loss_mask): `tokenized_prompt`, `tokenized_prompt_mask`, `tokenized_langact_mask`, ..., `token_loss_mask`."""
        fmt = self._resolve_format(is_vqa_sample, is_prediction_sample)
        formatted_prompt = fmt.format_prompt(
            prompt, state, state_type, time_horizon_seconds=time_horizon_seconds if not is_vqa_sample else None,
            frame_description=frame_description, state_dropout=state_dropout)
        pad_id = self._tokenizer.pad_id()
        tokens = self._tokenizer.encode(formatted_prompt, add_bos=True, add_eos=False)
        reasoning_start = len(tokens)
        if reasoning is not None:
            clean_reason = reasoning.strip().replace("_", " ").replace("\n", " ")
            tokens += self._tokenizer.encode(clean_reason, add_bos=False, add_eos=True)
        reasoning_end = len(tokens)
        if len(tokens) > self._max_len:
            tokens = tokens[: self._max_len]
            reasoning_end = min(reasoning_end, self._max_len)
        attn_mask, reasoning_mask, token_loss_mask = self._create_base_masks(len(tokens), reasoning_start, reasoning_end,
                                                                            reasoning is not None)
        if reasoning is None:
            number_mask = direction_mask = None
        else:
            token_loss_mask = self._apply_reasoning_dropout(token_loss_mask, reasoning_mask, is_vqa_sample)
            number_mask, direction_mask = self._build_number_direction_masks(tokens, reasoning_mask, fmt, is_vqa_sample)
        pad_count = self._max_len - len(tokens)
        if pad_count > 0:
            tokens = tokens + [pad_id] * pad_count
        return np.asarray(tokens, dtype=np.int32), attn_mask, reasoning_mask, number_mask, direction_mask, token_loss_mask

    def decode(self, tokens) -> str:
        """tokenizer.py:317-326."""
        if not isinstance(tokens, list):
            tokens = np.asarray(tokens).tolist()
        vocab_size = self._tokenizer.vocab_size()
        return self._tokenizer.decode([t for t in tokens if 0 <= t < vocab_size]).strip()

    def encode(self, text: str, add_bos: bool = False, add_eos: bool = False):
        return self._tokenizer.encode(text, add_bos=add_bos, add_eos=add_eos)
