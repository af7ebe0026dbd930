"""Drop-in ``GPT`` for the inference hot path, driven by liblitparrot_b200.so.

Public surface mirrors the reference's ``lit_gpt/model.py``: ``GPT(config)`` with the same module tree and
``state_dict()`` keys (model.py:24-36), ``forward(idx, max_seq_length=None, input_pos=None)`` (model.py:63-111)
with the same assertions, ``from_name``, ``reset_cache``, ``build_rope_cache`` / ``build_mask_cache`` /
``build_kv_caches``, ``_init_weights``, attributes ``config / rope_cache / mask_cache / kv_caches``.

What is different underneath (B200-first, not a translation):
  * the nn.Modules only *own* parameters; ``GPT.forward`` hands raw device pointers to hand-written sm_100a
    kernels through the C ABI (``_lib.py``) — there is no PyTorch compute on the path and no CPU fallback;
  * activations live in fp32 buffers; ``precision="bf16"`` rounds at the reference's bf16 rounding points,
    ``precision="fp32"`` (default) keeps fp32 activations over the stored (bf16 / int4 / ...) weights;
  * the KV cache is compact — one head per query group ``(B, G, max_seq, hs)`` instead of the reference's
    ``n_head`` replicated heads (model.py:130-144, 217-220) — and overflow is a ring index, not a roll;
  * a single-token step (T == 1 with ``input_pos``) is captured once into a CUDA graph and replayed.
"""
import math
from typing import Any, List, Optional, Tuple

import torch
import torch.nn as nn

from lit_parrot_b200 import _lib
from lit_parrot_b200.config import Config

RoPECache = Tuple[torch.Tensor, torch.Tensor]
KVCache = Tuple[torch.Tensor, torch.Tensor]

_DRIVEN = "lit_parrot_b200 sub-modules only own parameters; the computation is driven by GPT.forward"


class LayerNorm(torch.nn.LayerNorm):
    """Parameter container with torch.nn.LayerNorm's state-dict layout (weight, bias)."""


class RMSNorm(nn.Module):
    """Parameter container for lit_gpt/rmsnorm.py:4-21 (weight only)."""

    def __init__(self, size: int, dim: int = -1, eps: float = 1e-5) -> None:
        super().__init__()
        self.weight = nn.Parameter(torch.ones(size))
        self.eps = eps
        self.dim = dim

    def forward(self, x):  # pragma: no cover
        raise NotImplementedError(_DRIVEN)


class CausalSelfAttention(nn.Module):
    def __init__(self, config: Config) -> None:
        super().__init__()
        # `nn.Linear` is looked up at call time so that the `quantization()` plug-in switch applies (utils.py)
        # tensor parallel (config.with_tp): this rank's query groups / their heads only; tp_size == 1 is the reference layout
        self.attn = nn.Linear(config.n_embd, config.qkv_rows_local, bias=config.bias)
        self.proj = nn.Linear(config.n_head_local * config.head_size, config.n_embd, bias=config.bias)
        self.config = config

    def forward(self, *a, **k):  # pragma: no cover
        raise NotImplementedError(_DRIVEN)


class GptNeoxMLP(nn.Module):
    def __init__(self, config: Config) -> None:
        super().__init__()
        self.fc = nn.Linear(config.n_embd, config.intermediate_size_local, bias=config.bias)
        self.proj = nn.Linear(config.intermediate_size_local, config.n_embd, bias=config.bias)

    def forward(self, *a, **k):  # pragma: no cover
        raise NotImplementedError(_DRIVEN)


class LLaMAMLP(nn.Module):
    def __init__(self, config: Config) -> None:
        super().__init__()
        self.fc_1 = nn.Linear(config.n_embd, config.intermediate_size_local, bias=config.bias)
        self.fc_2 = nn.Linear(config.n_embd, config.intermediate_size_local, bias=config.bias)
        self.proj = nn.Linear(config.intermediate_size_local, config.n_embd, bias=config.bias)

    def forward(self, *a, **k):  # pragma: no cover
        raise NotImplementedError(_DRIVEN)


class Block(nn.Module):
    def __init__(self, config: Config) -> None:
        super().__init__()
        self.norm_1 = config.norm_class(config.n_embd, eps=config.norm_eps)
        self.attn = CausalSelfAttention(config)
        if not config.shared_attention_norm:
            self.norm_2 = config.norm_class(config.n_embd, eps=config.norm_eps)
        self.mlp = config.mlp_class(config)
        self.config = config

    def forward(self, *a, **k):  # pragma: no cover
        raise NotImplementedError(_DRIVEN)


def build_rope_cache(seq_len: int, n_elem: int, dtype: torch.dtype, device: torch.device, base: int = 10000,
                     condense_ratio: int = 1) -> RoPECache:
    """cos/sin tables of shape (seq_len, n_elem) — same construction and the same fp16 rounding rule for 16-bit
    working dtypes as the reference (model.py:304-327).  Host-side, one-off."""
    theta = 1.0 / (base ** (torch.arange(0, n_elem, 2, device=device) / n_elem))
    seq_idx = torch.arange(seq_len, device=device) / condense_ratio
    idx_theta = torch.outer(seq_idx, theta).repeat(1, 2)
    cos, sin = torch.cos(idx_theta), torch.sin(idx_theta)
    if dtype in (torch.float16, torch.bfloat16, torch.int8):
        return cos.half(), sin.half()
    return cos, sin


class GPT(nn.Module):
    def __init__(self, config: Config) -> None:
        super().__init__()
        assert config.padded_vocab_size is not None
        self.config = config
        self.lm_head = nn.Linear(config.n_embd, config.padded_vocab_size, bias=False)
        self.transformer = nn.ModuleDict(
            dict(
                wte=nn.Embedding(config.padded_vocab_size, config.n_embd),
                h=nn.ModuleList(self._make_block(config, i) for i in range(config.n_layer)),
                ln_f=config.norm_class(config.n_embd, eps=config.norm_eps),
            )
        )
        self.rope_cache: Optional[RoPECache] = None
        self.mask_cache: Optional[torch.Tensor] = None
        self.kv_caches: List[KVCache] = []
        # ---- B200 engine knobs -----------------------------------------------------------------
        self.precision = "fp32"          # "fp32": fp32 activations; "bf16": round where the reference's bf16-true rounds
        self.kv_cache_dtype: Optional[torch.dtype] = None  # default: parameter dtype (reference: model.py:236)
        self.use_cuda_graph = True
        self.tp_context = None           # lit_parrot_b200.tp.TPContext when config.tp_size > 1
        self._engine = None

    def _make_block(self, config: Config, block_idx: int) -> nn.Module:
        """Hook of the adapter / LoRA variants (lit_gpt/adapter.py:43, lit_gpt/lora.py:500): they build their own Block."""
        return Block(config)

    # ------------------------------------------------------------------ reference-compatible helpers
    def _init_weights(self, module: nn.Module) -> None:
        """model.py:41-54 (used through ``model.apply(model._init_weights)``)."""
        if isinstance(module, nn.Linear):
            torch.nn.init.normal_(module.weight, mean=0.0, std=0.02)
            if module.bias is not None:
                torch.nn.init.zeros_(module.bias)
        elif isinstance(module, nn.Embedding):
            torch.nn.init.normal_(module.weight, mean=0.0, std=0.02)
        elif isinstance(module, nn.LayerNorm):
            torch.nn.init.ones_(module.weight)
            torch.nn.init.zeros_(module.bias)
            module.eps = self.config.norm_eps
        elif isinstance(module, RMSNorm):
            torch.nn.init.ones_(module.weight)
            module.eps = self.config.norm_eps
        self._engine = None

    def reset_cache(self) -> None:
        self.kv_caches.clear()
        if self._engine is not None:
            self._engine.drop_graphs()

    @classmethod
    def from_name(cls, name: str, **kwargs: Any) -> "GPT":
        return cls(Config.from_name(name, **kwargs))

    def set_precision(self, precision: str) -> "GPT":
        if precision not in ("fp32", "bf16"):
            raise ValueError(f"precision must be 'fp32' or 'bf16', got {precision!r}")
        if precision != self.precision:
            self.precision = precision
            self.rope_cache = None
            self.invalidate()
        return self

    def invalidate(self) -> None:
        """Forget packed weights / captured graphs (call after editing parameters in place)."""
        self._engine = None

    def _apply(self, fn, *a, **k):  # .to() / .cuda() / .half() move storage: pointers go stale
        self._engine = None
        self.rope_cache = None
        self.kv_caches = []
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._engine = None
        return super().load_state_dict(*a, **k)

    def build_rope_cache(self, idx: torch.Tensor) -> RoPECache:
        dtype = torch.bfloat16 if self.precision == "bf16" else torch.float32
        return build_rope_cache(self.config.block_size, self.config.rope_n_elem, dtype, idx.device,
                                condense_ratio=self.config.condense_ratio)

    def build_mask_cache(self, idx: torch.Tensor) -> torch.Tensor:
        """Kept for API compatibility (model.py:126-128); the kernels never read a mask tensor."""
        bs = self.config.block_size
        return torch.tril(torch.ones((bs, bs), device=idx.device, dtype=torch.bool)).unsqueeze(0).unsqueeze(0)

    def _kv_dtype(self) -> torch.dtype:
        if self.kv_cache_dtype is not None:
            return self.kv_cache_dtype
        return self.transformer.wte.weight.dtype

    def build_kv_caches(self, idx: torch.Tensor, max_seq_length: int, rope_cache_length: int = 0) -> List[KVCache]:
        """Zero-filled caches, one (k, v) pair per layer, COMPACT layout (B, n_query_groups, max_seq, hs)."""
        cfg = self.config
        B = idx.size(0)
        shape = (cfg.n_layer, 2, B, cfg.n_query_groups_local, max_seq_length, cfg.head_size)
        store = torch.zeros(shape, device=idx.device, dtype=self._kv_dtype())
        return [(store[l, 0], store[l, 1]) for l in range(cfg.n_layer)]

    # ------------------------------------------------------------------ forward
    def forward(self, idx: torch.Tensor, max_seq_length: Optional[int] = None,
                input_pos: Optional[torch.Tensor] = None) -> torch.Tensor:
        return self._forward_impl(idx, max_seq_length, input_pos)

    def _forward_impl(self, idx: torch.Tensor, max_seq_length: Optional[int], input_pos: Optional[torch.Tensor],
                      last_only: bool = False, raw_logits: bool = False) -> torch.Tensor:
        """`last_only`: ln_f + lm_head on the last position only (what generate() consumes, base.py:136);
        `raw_logits`: return the engine's fp32 logits buffer instead of a copy in the parameter dtype."""
        B, T = idx.size()
        use_kv_cache = input_pos is not None
        block_size = self.config.block_size
        if max_seq_length is None:
            max_seq_length = block_size
        if use_kv_cache:
            assert max_seq_length >= T, f"Cannot forward sequence of length {T}, max seq length is only {max_seq_length}"
        assert max_seq_length <= block_size, f"Cannot attend to {max_seq_length}, block size is only {block_size}"
        assert block_size >= T, f"Cannot forward sequence of length {T}, block size is only {block_size}"
        if idx.device.type != "cuda":
            raise RuntimeError("lit_parrot_b200.GPT runs on CUDA (sm_100a) only; there is no CPU path — "
                               "move the model and inputs to a B200 (`model.cuda()`).")
        eng = self._get_engine(idx.device)
        if self.rope_cache is None:
            self.rope_cache = self.build_rope_cache(idx)
            eng.set_rope(self.rope_cache)
        if use_kv_cache:
            if not self.kv_caches:
                self.kv_caches = self.build_kv_caches(idx, max_seq_length)
                eng.drop_graphs()
            return eng.forward(idx, input_pos, self.kv_caches, allow_graph=not (last_only or raw_logits),
                               last_only=last_only, raw_logits=raw_logits)
        # no cache: causal attention over the T tokens through a scratch cache of length T
        scratch = eng.scratch_cache(B, T)
        pos = torch.arange(T, device=idx.device, dtype=torch.int32)
        return eng.forward(idx, pos, scratch, allow_graph=False)

    def _get_engine(self, device: torch.device):
        if self._engine is None or self._engine.device != device or self._engine.precision != self.precision:
            from lit_parrot_b200.engine import Engine

            self._engine = Engine(self, device)
            if self.rope_cache is not None:
                self._engine.set_rope(self.rope_cache)
        return self._engine
