"""Prompt tuning: trainable soft prompts living on the client (reference: src/petals/client/ptune.py:15-84).

``tuning_mode="ptune"`` prepends ``pre_seq_len`` learned embeddings to the input; ``"deep_ptune"``
additionally adds learned vectors to the first ``pre_seq_len`` positions of every block's input
(server side: csrc/elementwise.cu add_prompts_kernel / oracle add in server/backend.py).
Unlike the reference deep prompts are allocated for exactly ``num_hidden_layers``
blocks and documented as such; parameters are fp32 and created eagerly (no meta-device workaround)."""
from __future__ import annotations

import dataclasses
from typing import Optional, Tuple

import torch
import torch.nn as nn

from petals_b200.utils.misc import DUMMY

TUNING_MODES = (None, "ptune", "deep_ptune")


@dataclasses.dataclass
class PTuneConfig:
    pre_seq_len: int = 0  # number of learned prompt tokens
    tuning_mode: Optional[str] = None  # None | "ptune" | "deep_ptune"


class PTuneMixin:
    """Mixed into client model shells. Requires ``self.config`` and an embedding dtype."""

    def init_prompts(self, config) -> None:
        mode = getattr(config, "tuning_mode", None)
        if mode not in TUNING_MODES and not (mode and "ptune" in mode):
            raise NotImplementedError(f"tuning_mode={mode!r} is not supported (choose from {TUNING_MODES})")
        if mode and "ptune" in mode:
            if config.pre_seq_len <= 0:
                raise ValueError("pre_seq_len must be positive when prompt tuning is enabled")
            self.pre_seq_len = config.pre_seq_len
            self.prefix_tokens = torch.arange(self.pre_seq_len).long()
            self.prompt_embeddings = nn.Embedding(self.pre_seq_len, config.hidden_size, dtype=torch.float32)
            if mode == "deep_ptune":
                self.intermediate_prompt_embeddings = nn.Embedding(
                    self.pre_seq_len, config.num_hidden_layers * config.hidden_size, dtype=torch.float32)
                nn.init.zeros_(self.intermediate_prompt_embeddings.weight)  # start as a no-op perturbation
        else:
            self.pre_seq_len = 0

    def get_prompt(self, batch_size: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """Returns (prompts [B, pre, H], deep prompts [L, B, pre, H] or DUMMY) in the embedding dtype."""
        dev = self.prompt_embeddings.weight.device
        tokens = self.prefix_tokens.to(dev).unsqueeze(0).expand(batch_size, -1)
        prompts = self.prompt_embeddings(tokens)
        deep = DUMMY
        if self.config.tuning_mode == "deep_ptune":
            deep = self.intermediate_prompt_embeddings(tokens)
            deep = deep.view(batch_size, self.pre_seq_len, self.config.num_hidden_layers, self.config.hidden_size)
            deep = deep.permute(2, 0, 1, 3)
        dtype = self.word_embeddings_dtype()
        return prompts.to(dtype), (deep.to(dtype) if deep is not DUMMY else deep)

    def word_embeddings_dtype(self) -> torch.dtype:
        return self.get_input_embeddings().weight.dtype
