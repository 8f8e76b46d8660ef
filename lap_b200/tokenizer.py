"""Token / mask producer of the lang-action prompt (second slice).  Behaviour = the reference's
`PaligemmaTokenizer.tokenize` (src/lap/models/tokenizer.py:237-315, with the mask rules of :105-207 and `is_number` of
src/lap/models/prompt_utils/checkers.py:4-6), restated as ONE span computation: a sample is (n_prompt, n_total) and every
mask is `arange(max_len)` compared against that span.

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
    """One tokenised sample is fully described by two integers — `n_prompt` (BOS + formatted prompt) and `n_total`
    (… + cleaned reasoning + EOS, cut at `max_len`) — so every mask is a comparison of `arange(max_len)` against that
    span; the per-token number / direction classes come from a memoised id -> (is_number, is_direction) table."""

    def __init__(self, sp_processor, max_len: int = 48, prompt_format="lap", prediction_format="default",
                 reasoning_mask_prob: float = 0.0):
        """tokenizer.py:221-235 with the processor passed in (`sentencepiece.SentencePieceProcessor(model_file=...)`)."""
        self._tokenizer = sp_processor
        self._max_len = int(max_len)
        self._formats = {
            "action": _resolve(prompt_format, PROMPT_FORMAT_REGISTRY, "prompt format"),
            "prediction": _resolve(prediction_format, PREDICTION_PROMPT_FORMAT_REGISTRY, "prediction format"),
            "vqa": DEFAULT_VQA_PROMPT_FORMAT,
        }
        self.reasoning_mask_prob = float(reasoning_mask_prob)
        self._slots = np.arange(self._max_len)
        self._piece_class: dict[tuple[int, int], tuple[bool, bool]] = {}

    # kept as attributes for callers that introspect the tokenizer the way they do the reference's
    @property
    def _prompt_format(self):
        return self._formats["action"]

    @property
    def _prediction_format(self):
        return self._formats["prediction"]

    @property
    def _vqa_format(self):
        return self._formats["vqa"]

    def _classes(self, ids: np.ndarray, fmt) -> tuple[np.ndarray, np.ndarray]:
        """(is_number, is_direction) of each token id (tokenizer.py:174-207: classified on the SentencePiece piece; an
        empty piece is neither).  Memoised per (format, id)."""
        key0 = id(fmt)
        num = np.zeros(len(ids), dtype=bool)
        dirn = np.zeros(len(ids), dtype=bool)
        for n, tid in enumerate(ids.tolist()):
            hit = self._piece_class.get((key0, tid))
            if hit is None:
                piece = self._tokenizer.id_to_piece(tid)
                hit = (bool(piece) and bool(is_number(piece)), bool(piece) and bool(fmt.direction_token_checker(piece)))
                self._piece_class[(key0, tid)] = hit
            num[n], dirn[n] = hit
        return num, dirn

    def tokenize(self, prompt: str, reasoning: str | None = None, state=None, state_type: str | None = None, *,
                 is_vqa_sample: bool = False, is_prediction_sample: bool = False,
                 time_horizon_seconds: float | None = None, frame_description: str = "robot base frame",
                 state_dropout: float = 0.0):
        """tokenizer.py:237-315 -> (tokens int32[max_len], attn_mask, reasoning_mask | None, number_mask | None,
        direction_mask | None, token_loss_mask), i.e. `tokenized_prompt`, `tokenized_prompt_mask`,
        `tokenized_langact_mask`, the two metric masks and `token_loss_mask`."""
        L, sp = self._max_len, self._tokenizer
        # which format (tokenizer.py:93-103: prediction wins over vqa)
        fmt = self._formats["prediction" if is_prediction_sample else "vqa" if is_vqa_sample else "action"]
        text = fmt.format_prompt(prompt, state, state_type,
                                 time_horizon_seconds=None if is_vqa_sample else time_horizon_seconds,
                                 frame_description=frame_description, state_dropout=state_dropout)
        ids = sp.encode(text, add_bos=True, add_eos=False)
        n_prompt = len(ids)
        if reasoning is not None:
            ids = ids + sp.encode(reasoning.strip().replace("_", " ").replace("\n", " "), add_bos=False, add_eos=True)
        n_total = min(len(ids), L)                       # truncation (:265-269)
        tokens = np.full(L, sp.pad_id(), dtype=np.int32)  # right padding (:305-311)
        tokens[:n_total] = ids[:n_total]
        attn_mask = self._slots < n_total
        token_loss_mask = np.ones(L, dtype=bool)
        if reasoning is None:                            # (:105-138: no reasoning span -> no span masks)
            return tokens, attn_mask, None, None, None, token_loss_mask
        span = (self._slots >= min(n_prompt, L)) & (self._slots < n_total)
        where = np.flatnonzero(span)
        p = self.reasoning_mask_prob
        if not 0.0 <= p <= 1.0:
            raise ValueError(f"reasoning_mask_prob must be between 0.0 and 1.0, got {p}")
        if p > 0.0 and not is_vqa_sample and where.size:
            # (:140-172) one uniform draw per reasoning token from numpy's GLOBAL stream, as the reference does
            token_loss_mask[where[np.random.rand(where.size) < p]] = False
        number_mask = np.zeros(L, dtype=bool)
        direction_mask = np.zeros(L, dtype=bool)
        if not is_vqa_sample and where.size:
            number_mask[where], direction_mask[where] = self._classes(tokens[where], fmt)
        return tokens, attn_mask, span, number_mask, direction_mask, token_loss_mask

    def decode(self, tokens) -> str:
        """tokenizer.py:317-326: ids outside the vocabulary are dropped before detokenising."""
        ids = np.asarray(tokens).reshape(-1)
        keep = ids[(ids >= 0) & (ids < self._tokenizer.vocab_size())]
        return self._tokenizer.decode(keep.tolist()).strip()

    def encode(self, text: str, add_bos: bool = False, add_eos: bool = False):
        return self._tokenizer.encode(text, add_bos=add_bos, add_eos=add_eos)
