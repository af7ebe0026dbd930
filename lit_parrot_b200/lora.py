"""Merged-LoRA inference — drop-in surface of the reference's ``lit_gpt/lora.py``.

Same ``Config`` fields (lora.py:449-476), module tree and ``state_dict()`` keys (``...attn.attn.lora_A`` / ``lora_B``, optional
LoRA on ``attn.proj``, the MLP and ``lm_head``; ``lora_ind`` is a non-persistent buffer as in lora.py:294), and the same entry
point ``merge_lora_weights(model)`` (lora.py:676-680) that ``generate/lora.py:117`` calls before generating.  What is different:

  * the modules only own parameters; the merge ``W += (B @ A) * alpha / r`` — for the fused QKV matrix the grouped form
    ``conv1d(A, B, groups = #enabled)`` scattered to the rows of ``lora_ind`` (lora.py:338-361) — runs on the device as
    ``lp_lora_merge`` (csrc/adapter.cu) directly on the stored weights (fp32 or bf16, rounded like ``weight.data += ...``);
  * there is no un-merged forward: after the merge the model IS a plain GPT and runs through the same kernels (persistent step
    kernel included).  ``GPT.forward`` on a model with pending LoRA updates raises.
"""
import math
from dataclasses import dataclass
from typing import Any, Optional, Tuple, Union

import torch
import torch.nn as nn

from lit_parrot_b200 import _lib
from lit_parrot_b200 import model as base
from lit_parrot_b200.config import Config as BaseConfig


def _merge_rows(weight: torch.Tensor, B_rows: torch.Tensor, A: torch.Tensor, rows: Optional[torch.Tensor], scaling: float) -> None:
    """weight[rows] += scaling * (B_rows @ A) on the device (lp_lora_merge)."""
    if weight.device.type != "cuda":
        raise RuntimeError("lit_parrot_b200 merges LoRA weights on the GPU (lp_lora_merge); move the model to CUDA first — "
                           "there is no CPU path")
    if weight.dtype not in (torch.float32, torch.bfloat16):
        raise RuntimeError(f"LoRA merge needs float32 / bfloat16 weights, got {weight.dtype}")
    if not weight.is_contiguous():  # fc_1 / fc_2 of an engine that already interleaved them: merge a dense copy, write it back
        dense = weight.contiguous()
        _merge_rows(dense, B_rows, A, rows, scaling)
        weight.copy_(dense)
        return
    lib = _lib.init(weight.device.index if weight.device.index is not None else torch.cuda.current_device())
    Bf, Af = B_rows.detach().float().contiguous(), A.detach().float().contiguous()
    ri = None if rows is None else rows.to(device=weight.device, dtype=torch.int32).contiguous()
    n_rows, r = Bf.shape
    stream = torch.cuda.current_stream(weight.device).cuda_stream
    _lib.check(lib.lp_lora_merge(weight.data_ptr(), _lib.LP_BF16 if weight.dtype == torch.bfloat16 else _lib.LP_F32, weight.shape[0],
                                 weight.shape[1], Bf.data_ptr(), Af.data_ptr(), r, None if ri is None else ri.data_ptr(), n_rows,
                                 float(scaling), stream), "lp_lora_merge")
    torch.cuda.current_stream(weight.device).synchronize()  # Bf / Af / ri are temporaries


class LoRALayer:
    """Attribute holder of lora.py:64-88."""

    def __init__(self, r: int, lora_alpha: int, lora_dropout: float, merge_weights: bool):
        self.r = r
        self.lora_alpha = lora_alpha
        self.lora_dropout = nn.Dropout(p=lora_dropout) if lora_dropout > 0.0 else (lambda x: x)
        self.merged = False
        self.merge_weights = merge_weights


class LoRALinear(nn.Linear, LoRALayer):
    """nn.Linear parameters + ``lora_A`` (r, in) / ``lora_B`` (out, r) (lora.py:91-143)."""

    def __init__(self, in_features: int, out_features: int, r: int = 0, lora_alpha: int = 1, lora_dropout: float = 0.0,
                 fan_in_fan_out: bool = False, merge_weights: bool = True, **kwargs):
        super().__init__(in_features, out_features, **kwargs)
        LoRALayer.__init__(self, r=r, lora_alpha=lora_alpha, lora_dropout=lora_dropout, merge_weights=merge_weights)
        if fan_in_fan_out:
            raise NotImplementedError("fan_in_fan_out storage is not used by any Lit-GPT model")
        self.fan_in_fan_out = False
        if r > 0:
            self.lora_A = nn.Parameter(self.weight.new_zeros((r, in_features)))
            self.lora_B = nn.Parameter(self.weight.new_zeros((out_features, r)))
            self.scaling = self.lora_alpha / self.r
            self.weight.requires_grad = False
        self.reset_parameters()

    def reset_parameters(self):
        super().reset_parameters()
        if hasattr(self, "lora_A"):
            nn.init.kaiming_uniform_(self.lora_A, a=math.sqrt(5))
            nn.init.zeros_(self.lora_B)

    def has_pending_update(self) -> bool:
        return self.r > 0 and hasattr(self, "lora_A") and not self.merged

    def merge(self):
        """W = W + (lora_B @ lora_A) * scaling, once (lora.py:154-164)."""
        if self.merge_weights and not self.merged:
            if self.r > 0:
                _merge_rows(self.weight.data, self.lora_B.data, self.lora_A.data, None, self.scaling)
            self.merged = True

    def forward(self, x):  # pragma: no cover
        raise NotImplementedError(base._DRIVEN)


class LoRAQKVLinear(LoRALinear):
    """The fused QKV projection with LoRA on any of q / k / v (lora.py:182-294): ``lora_A`` (r * #enabled, in), ``lora_B``
    (rows of the enabled parts, r), ``lora_ind`` = the rows of ``weight`` the update goes to."""

    def __init__(self, in_features: int, out_features: int, n_head: int, n_query_groups: int, r: int = 0, lora_alpha: int = 1,
                 lora_dropout: float = 0.0, enable_lora: Union[bool, Tuple[bool, bool, bool]] = False, fan_in_fan_out: bool = False,
                 merge_weights: bool = True, **kwargs):
        super().__init__(in_features, out_features, **kwargs)  # r = 0: the parent allocates nothing
        LoRALayer.__init__(self, r=r, lora_alpha=lora_alpha, lora_dropout=lora_dropout, merge_weights=merge_weights)
        if isinstance(enable_lora, bool):
            enable_lora = [enable_lora] * 3
        assert len(enable_lora) == 3
        if fan_in_fan_out:
            raise NotImplementedError("fan_in_fan_out storage is not used by any Lit-GPT model")
        self.enable_lora = enable_lora
        if r > 0 and any(enable_lora):
            enable_q, enable_k, enable_v = enable_lora
            self.lora_A = nn.Parameter(self.weight.new_zeros((r * sum(enable_lora), in_features)))
            self.kv_embd_size = self.in_features // (n_head // n_query_groups)
            rows = self.in_features * enable_q + self.kv_embd_size * enable_k + self.kv_embd_size * enable_v
            self.lora_B = nn.Parameter(self.weight.new_zeros(rows, r))
            self.scaling = self.lora_alpha / self.r
            self.weight.requires_grad = False
            dev = self.weight.device
            E, kv = self.in_features, self.kv_embd_size
            spans = [(0, E)] * enable_q + [(E, E + kv)] * enable_k + [(E + kv, self.out_features)] * enable_v
            self.register_buffer("lora_ind", torch.cat([torch.arange(a, b, device=dev) for a, b in spans]), persistent=False)
        self.reset_parameters()

    def has_pending_update(self) -> bool:
        return self.r > 0 and any(self.enable_lora) and not self.merged

    def merge(self):
        """delta = conv1d(lora_A, lora_B, groups = #enabled) * scaling, zero-padded to the rows of lora_ind (lora.py:338-361): the
        grouped convolution cuts lora_B's rows into #enabled EQUAL blocks, block j multiplies rows [j r, (j+1) r) of lora_A."""
        if self.merge_weights and not self.merged:
            if self.r > 0 and any(self.enable_lora):
                ng = sum(self.enable_lora)
                rows = self.lora_B.shape[0]
                if rows % ng:
                    raise RuntimeError(f"conv1d: {rows} LoRA rows are not divisible by {ng} groups (the reference fails here too)")
                blk = rows // ng
                for j in range(ng):
                    _merge_rows(self.weight.data, self.lora_B.data[j * blk:(j + 1) * blk], self.lora_A.data[j * self.r:(j + 1) * self.r],
                                self.lora_ind[j * blk:(j + 1) * blk], self.scaling)
            self.merged = True


def mark_only_lora_as_trainable(model: nn.Module, bias: str = "none") -> None:
    """lora.py:412-442."""
    if bias not in ("none", "all", "lora_only"):
        raise NotImplementedError
    for n, p in model.named_parameters():
        if "lora_" not in n:
            p.requires_grad = False
    if bias == "all":
        for n, p in model.named_parameters():
            if "bias" in n:
                p.requires_grad = True
    elif bias == "lora_only":
        for m in model.modules():
            if isinstance(m, LoRALayer) and getattr(m, "bias", None) is not None:
                m.bias.requires_grad = True


def lora_filter(key: str, value: Any) -> bool:
    return "lora_" in key


@dataclass
class Config(BaseConfig):
    """lora.py:449-476: rank, alpha, dropout and which projections carry LoRA."""

    r: int = 0.0
    alpha: int = 1.0
    dropout: float = 0.0
    to_query: bool = False
    to_key: bool = False
    to_value: bool = False
    to_projection: bool = False
    to_mlp: bool = False
    to_head: bool = False

    @property
    def mlp_class(self):
        import lit_parrot_b200.lora as me

        return getattr(me if self.to_mlp else base, self._mlp_class)


def _lora_linear(config: Config, fan_in: int, fan_out: int, enabled: bool, bias: bool) -> nn.Module:
    if not enabled:
        return nn.Linear(fan_in, fan_out, bias=bias)
    return LoRALinear(fan_in, fan_out, bias=bias, r=config.r, lora_alpha=config.alpha, lora_dropout=config.dropout)


class CausalSelfAttention(base.CausalSelfAttention):
    def __init__(self, config: Config) -> None:
        nn.Module.__init__(self)
        self.attn = LoRAQKVLinear(in_features=config.n_embd, out_features=config.qkv_rows, r=config.r, lora_alpha=config.alpha,
                                  lora_dropout=config.dropout, enable_lora=(config.to_query, config.to_key, config.to_value),
                                  bias=config.bias, n_head=config.n_head, n_query_groups=config.n_query_groups)
        self.proj = _lora_linear(config, config.n_embd, config.n_embd, config.to_projection, config.bias)
        self.config = config


class GptNeoxMLP(base.GptNeoxMLP):
    def __init__(self, config: Config) -> None:
        nn.Module.__init__(self)
        self.fc = _lora_linear(config, config.n_embd, config.intermediate_size, True, config.bias)
        self.proj = _lora_linear(config, config.intermediate_size, config.n_embd, True, config.bias)


class LLaMAMLP(base.LLaMAMLP):
    def __init__(self, config: Config) -> None:
        nn.Module.__init__(self)
        self.fc_1 = _lora_linear(config, config.n_embd, config.intermediate_size, True, config.bias)
        self.fc_2 = _lora_linear(config, config.n_embd, config.intermediate_size, True, config.bias)
        self.proj = _lora_linear(config, config.intermediate_size, config.n_embd, True, config.bias)


class Block(base.Block):
    def __init__(self, config: Config) -> None:
        nn.Module.__init__(self)
        self.norm_1 = config.norm_class(config.n_embd, eps=config.norm_eps)
        self.attn = CausalSelfAttention(config)
        if not config.shared_attention_norm:
            self.norm_2 = config.norm_class(config.n_embd, eps=config.norm_eps)
        self.mlp = config.mlp_class(config)
        self.config = config


class GPT(base.GPT):
    def __init__(self, config: Config) -> None:
        if config.tp_size > 1:
            raise NotImplementedError("merge the LoRA weights on one device, then shard the merged state dict (tp.shard_state_dict)")
        super().__init__(config)
        if config.to_head:
            self.lm_head = _lora_linear(config, config.n_embd, config.padded_vocab_size, True, False)

    def _make_block(self, config: Config, block_idx: int) -> nn.Module:
        return Block(config)

    def forward(self, idx: torch.Tensor, max_seq_length: Optional[int] = None, input_pos: Optional[torch.Tensor] = None,
                lm_head_chunk_size: int = 0):
        pending = [n for n, m in self.named_modules() if isinstance(m, LoRALinear) and m.has_pending_update()]
        if pending:
            raise RuntimeError(f"{len(pending)} LoRA layers (first: {pending[0]!r}) are not merged: lit_parrot_b200 runs merged-LoRA "
                               "inference — call merge_lora_weights(model) after loading (generate/lora.py:117)")
        logits = self._forward_impl(idx, max_seq_length, input_pos)
        if lm_head_chunk_size > 0:
            return list(logits.split(lm_head_chunk_size, dim=1))
        return logits

    @classmethod
    def from_name(cls, name: str, **kwargs: Any) -> "GPT":
        return cls(Config.from_name(name, **kwargs))


def merge_lora_weights(model: GPT) -> None:
    """Merge LoRA weights into the full-rank weights (lora.py:676-680); the packed weight records of an existing engine are
    rebuilt on the next forward."""
    for module in model.modules():
        if isinstance(module, LoRALinear):
            module.merge()
    if hasattr(model, "invalidate"):
        model.invalidate()
