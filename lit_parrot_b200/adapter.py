"""LLaMA-Adapter inference — drop-in surface of the reference's ``lit_gpt/adapter.py``.

``Config`` (adapter.py:25-28) adds ``adapter_prompt_length`` / ``adapter_start_layer``; every attention layer with
``block_idx >= adapter_start_layer`` owns ``adapter_wte`` (aT, n_embd) and ``gating_factor`` (1, n_head, 1, 1) — same module
tree and ``state_dict()`` keys as the reference (adapter.py:167-177).  The arithmetic of adapter.py:234-254,

    y = y + gating_factor * softmax(q . ak^T / sqrt(hs)) . av,    (ak, av) = k / v parts of attn.attn(adapter_wte.weight)

is ``lp_adapter_attn`` (csrc/adapter.cu), launched behind the causal attention kernel of every adapted layer; the prefix
keys / values do not depend on the input and are computed once per engine (the reference's ``adapter_kv_caches``).
Adapter models decode through the per-op kernels (5 launches per layer + this one), not the persistent step kernel.
"""
from dataclasses import dataclass
from typing import Any, List, Optional

import torch
import torch.nn as nn

from lit_parrot_b200 import model as base
from lit_parrot_b200.config import Config as BaseConfig
from lit_parrot_b200.model import KVCache


@dataclass
class Config(BaseConfig):
    adapter_prompt_length: int = 10
    adapter_start_layer: int = 2


class CausalSelfAttention(base.CausalSelfAttention):
    """Parameter container of adapter.py:164-177."""

    def __init__(self, config: Config, block_idx: int) -> None:
        super().__init__(config)
        if block_idx >= config.adapter_start_layer:
            self.adapter_wte = nn.Embedding(config.adapter_prompt_length, config.n_embd)  # adapter embedding layer
            self.gating_factor = torch.nn.Parameter(torch.zeros(1, config.n_head, 1, 1))  # gate for adaption
        self.block_idx = block_idx


class Block(base.Block):
    def __init__(self, config: Config, block_idx: int) -> None:
        nn.Module.__init__(self)
        self.norm_1 = config.norm_class(config.n_embd, eps=config.norm_eps)
        self.attn = CausalSelfAttention(config, block_idx)
        if not config.shared_attention_norm:
            self.norm_2 = config.norm_class(config.n_embd, eps=config.norm_eps)
        self.mlp = config.mlp_class(config)
        self.config = config


class GPT(base.GPT):
    def __init__(self, config: Config) -> None:
        if config.tp_size > 1:
            raise NotImplementedError("adapter models are not sharded tensor-parallel")
        super().__init__(config)
        self.adapter_kv_caches: List[KVCache] = []

    def _make_block(self, config: Config, block_idx: int) -> nn.Module:
        return Block(config, block_idx)

    def reset_cache(self) -> None:
        super().reset_cache()
        self.adapter_kv_caches.clear()
        if self._engine is not None:
            self._engine.drop_adapter_kv()

    def forward(self, idx: torch.Tensor, max_seq_length: Optional[int] = None, input_pos: Optional[torch.Tensor] = None,
                lm_head_chunk_size: int = 0):
        logits = self._forward_impl(idx, max_seq_length, input_pos)
        if input_pos is not None and not self.adapter_kv_caches:
            # the reference fills `adapter_kv_caches` on the first cached forward (adapter.py:103-107); here they are views of the
            # engine's per-layer prefix keys / values, compact (1, G, aT, hs) like the KV cache (None below adapter_start_layer)
            self.adapter_kv_caches = self._engine.adapter_kv_views()
        if lm_head_chunk_size > 0:  # adapter.py:111-113 (a training-memory knob: same values, returned in chunks)
            return list(logits.split(lm_head_chunk_size, dim=1))
        return logits

    @classmethod
    def from_name(cls, name: str, **kwargs: Any) -> "GPT":
        return cls(Config.from_name(name, **kwargs))


def mark_only_adapter_as_trainable(model: GPT) -> None:
    """Sets `requires_grad=False` for all non-adapter weights (adapter.py:262-265)."""
    for name, param in model.named_parameters():
        param.requires_grad = adapter_filter(name, param)


def adapter_filter(key: str, value: Any) -> bool:
    return "adapter_wte" in key or "gating_factor" in key
