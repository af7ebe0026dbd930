"""Hyper-parameter presets for the inference hot path.

Mirrors the public surface of the reference's ``lit_gpt/config.py`` (``Config`` dataclass at
config.py:11-92, ``name_to_config`` at config.py:528): same field names, same derived values
(``padded_vocab_size`` config.py:56-58, ``n_query_groups`` config.py:60-63, ``intermediate_size``
config.py:65-68, ``head_size`` config.py:70-72).  The presets are *data*; they are generated here
from compact per-family tables rather than spelled out dict by dict.
"""
from dataclasses import dataclass
from typing import Any, Dict, Optional


def find_multiple(n: int, k: int) -> int:
    """Smallest multiple of ``k`` that is >= ``n`` (reference: lit_gpt/utils.py:19-23)."""
    assert k > 0
    return n if n % k == 0 else n + k - (n % k)


@dataclass
class Config:
    org: str = "Lightning-AI"
    name: str = "lit-GPT"
    block_size: int = 4096
    vocab_size: int = 50254
    padding_multiple: int = 512
    padded_vocab_size: Optional[int] = None
    n_layer: int = 16
    n_head: int = 32
    n_embd: int = 4096
    rotary_percentage: float = 0.25
    parallel_residual: bool = True
    bias: bool = True
    # n_query_groups == n_head -> MHA, == 1 -> MQA, in between -> GQA
    n_query_groups: Optional[int] = None
    shared_attention_norm: bool = False
    _norm_class: str = "LayerNorm"  # "LayerNorm" | "RMSNorm"
    norm_eps: float = 1e-5
    _mlp_class: str = "GptNeoxMLP"  # "GptNeoxMLP" | "LLaMAMLP"
    intermediate_size: Optional[int] = None
    condense_ratio: int = 1

    def __post_init__(self) -> None:
        assert self.n_embd % self.n_head == 0
        if self.padded_vocab_size is None:
            self.padded_vocab_size = find_multiple(self.vocab_size, self.padding_multiple)
        if self.n_query_groups is None:
            self.n_query_groups = self.n_head
        else:
            assert self.n_head % self.n_query_groups == 0
        if self.intermediate_size is None:
            if self._mlp_class == "LLaMAMLP":
                raise ValueError("The config needs to set the `intermediate_size`")
            self.intermediate_size = 4 * self.n_embd

    # ---- derived quantities -------------------------------------------------------------
    @property
    def head_size(self) -> int:
        return self.n_embd // self.n_head

    @property
    def rope_n_elem(self) -> int:
        """Number of leading head dims that are rotated (model.py:120, 222)."""
        return int(self.rotary_percentage * self.head_size)

    @property
    def q_per_kv(self) -> int:
        return self.n_head // self.n_query_groups

    @property
    def qkv_rows(self) -> int:
        """Output rows of the fused QKV projection (model.py:186)."""
        return (self.n_head + 2 * self.n_query_groups) * self.head_size

    # ---- tensor-parallel view (not part of the reference's Config: plain attributes, not dataclass fields) ----
    @property
    def tp_size(self) -> int:
        return getattr(self, "_tp_size", 1)

    @property
    def tp_rank(self) -> int:
        return getattr(self, "_tp_rank", 0)

    def with_tp(self, tp_size: int, tp_rank: int) -> "Config":
        """A copy of this config whose `*_local` quantities describe rank `tp_rank`'s shard: query groups (with their q
        heads) and MLP columns are split evenly; everything else is replicated."""
        import copy

        if tp_size < 1 or not 0 <= tp_rank < tp_size:
            raise ValueError(f"bad tensor-parallel coordinates {tp_rank}/{tp_size}")
        if self.n_query_groups % tp_size or self.intermediate_size % tp_size:
            raise ValueError(f"n_query_groups={self.n_query_groups} and intermediate_size={self.intermediate_size} must be "
                             f"divisible by the tensor-parallel size {tp_size} (query groups are the sharding unit)")
        c = copy.copy(self)
        object.__setattr__(c, "_tp_size", tp_size)
        object.__setattr__(c, "_tp_rank", tp_rank)
        return c

    @property
    def n_head_local(self) -> int:
        return self.n_head // self.tp_size

    @property
    def n_query_groups_local(self) -> int:
        return self.n_query_groups // self.tp_size

    @property
    def intermediate_size_local(self) -> int:
        return self.intermediate_size // self.tp_size

    @property
    def qkv_rows_local(self) -> int:
        return self.qkv_rows // self.tp_size

    @classmethod
    def from_name(cls, name: str, **kwargs: Any) -> "Config":
        conf = dict(name_to_config[name])
        conf.update(kwargs)
        return cls(**conf)

    @property
    def mlp_class(self):
        from lit_parrot_b200 import model

        return getattr(model, self._mlp_class)

    @property
    def norm_class(self):
        from lit_parrot_b200 import model

        return model.RMSNorm if self._norm_class == "RMSNorm" else model.LayerNorm


# --------------------------------------------------------------------------------------------
# Preset tables.  Columns are per family; every row expands to one (or several) named presets.
# --------------------------------------------------------------------------------------------
name_to_config: Dict[str, Dict[str, Any]] = {}


def _register(**kw: Any) -> None:
    name_to_config[kw["name"]] = kw


# Stability AI StableLM-alpha (config.py:98-107): NeoX defaults of the dataclass.
for _nm, _extra in (
    ("stablelm-base-alpha-3b", dict(padding_multiple=512)),
    ("stablelm-base-alpha-7b", dict(n_head=48, n_embd=6144, padding_multiple=256)),
    ("stablelm-tuned-alpha-3b", dict(n_head=32, padding_multiple=512)),
    ("stablelm-tuned-alpha-7b", dict(n_head=48, n_embd=6144, padding_multiple=256)),
):
    _register(org="stabilityai", name=_nm, **_extra)

# EleutherAI Pythia (config.py:112-146): (suffix, n_layer, n_embd, n_head, padding_multiple)
_PYTHIA = (
    ("70m", 6, 512, 8, 128),
    ("160m", 12, 768, 12, 128),
    ("410m", 24, 1024, 16, 128),
    ("1b", 16, 2048, 8, 128),
    ("1.4b", 24, 2048, 16, 128),
    ("2.8b", 32, 2560, 32, 128),
    ("6.9b", 32, 4096, 32, 256),
    ("12b", 36, 5120, 40, 512),
)
for _dedup in ("", "-deduped"):
    for _sz, _L, _E, _H, _pm in _PYTHIA:
        _register(
            org="EleutherAI", name=f"pythia-{_sz}{_dedup}", block_size=2048, n_layer=_L, n_embd=_E, n_head=_H,
            padding_multiple=_pm,
        )

# togethercomputer RedPajama-INCITE (config.py:152-194): NeoX, sequential residual, full rotary.
for _pattern, _E in (("RedPajama-INCITE-{}-3B-v1", 2560), ("RedPajama-INCITE-7B-{}", 4096),
                     ("RedPajama-INCITE-{}-7B-v0.1", 4096)):
    for _kind in ("Base", "Chat", "Instruct"):
        _register(
            org="togethercomputer", name=_pattern.format(_kind), block_size=2048, n_layer=32, n_embd=_E, n_head=32,
            padding_multiple=256, rotary_percentage=1.0, parallel_residual=False,
        )

# TII Falcon (config.py:200-236): parallel residual, no biases; 7b = MQA + one shared norm.
for _kind in ("", "-instruct"):
    _register(
        org="tiiuae", name=f"falcon-7b{_kind}", block_size=2048, padded_vocab_size=65024, n_layer=32, n_head=71,
        n_embd=4544, rotary_percentage=1.0, parallel_residual=True, n_query_groups=1, bias=False,
        shared_attention_norm=True,
    )
    _register(
        org="tiiuae", name=f"falcon-40b{_kind}", block_size=2048, padded_vocab_size=65024, n_layer=60, n_head=128,
        n_embd=8192, rotary_percentage=1.0, parallel_residual=True, n_query_groups=8, bias=False,
    )


def _llama_family(org: str, name: str, n_layer: int, n_head: int, n_embd: int, intermediate_size: int,
                  block_size: int, norm_eps: float, **extra: Any) -> None:
    """LLaMA-style decoder: RMSNorm, SwiGLU MLP, sequential residual, full rotary, no biases."""
    kw: Dict[str, Any] = dict(
        org=org, name=name, block_size=block_size, vocab_size=32000, padding_multiple=64, n_layer=n_layer,
        n_head=n_head, n_embd=n_embd, rotary_percentage=1.0, parallel_residual=False, bias=False,
        _norm_class="RMSNorm", norm_eps=norm_eps, _mlp_class="LLaMAMLP", intermediate_size=intermediate_size,
    )
    kw.update(extra)
    _register(**kw)


# (n_layer, n_head, n_embd, intermediate_size) by nominal size
_LLAMA_DIMS = {"3b": (26, 32, 3200, 8640), "7b": (32, 32, 4096, 11008), "13b": (40, 40, 5120, 13824),
               "33b": (60, 52, 6656, 17920), "70b": (80, 64, 8192, 28672)}

# OpenLM Research Open LLaMA (config.py:242-298)
for _sz in ("3b", "7b", "13b"):
    _llama_family("openlm-research", f"open_llama_{_sz}", *_LLAMA_DIMS[_sz], block_size=2048, norm_eps=1e-6)
# LMSYS Vicuna (config.py:304-360)
for _sz in ("7b", "13b", "33b"):
    _llama_family("lmsys", f"vicuna-{_sz}-v1.3", *_LLAMA_DIMS[_sz], block_size=2048, norm_eps=1e-6)
# LMSYS LongChat (config.py:366-406): 16k context through position interpolation.
for _sz in ("7b", "13b"):
    _llama_family("lmsys", f"longchat-{_sz}-16k", *_LLAMA_DIMS[_sz], block_size=16384, norm_eps=1e-6,
                  condense_ratio=8)
# NousResearch Hermes (config.py:412-431): explicit padded vocab, no vocab_size/padding_multiple keys.
_llama_family("NousResearch", "Nous-Hermes-13b", *_LLAMA_DIMS["13b"], block_size=2048, norm_eps=1e-6,
              padded_vocab_size=32001)
for _k in ("vocab_size", "padding_multiple"):
    del name_to_config["Nous-Hermes-13b"][_k]
# Meta Llama 2 (config.py:437-498); 70b uses 8 query groups.
for _kind in ("", "-chat"):
    for _sz in ("7b", "13b"):
        _llama_family("meta-llama", f"Llama-2-{_sz}{_kind}-hf", *_LLAMA_DIMS[_sz], block_size=4096, norm_eps=1e-5)
    _llama_family("meta-llama", f"Llama-2-70b{_kind}-hf", *_LLAMA_DIMS["70b"], block_size=4096, norm_eps=1e-5,
                  n_query_groups=8)
# Stability AI FreeWilly2 (config.py:504-525)
_llama_family("stabilityai", "FreeWilly2", *_LLAMA_DIMS["70b"], block_size=4096, norm_eps=1e-5, n_query_groups=8)

configs = list(name_to_config.values())
