"""`lit_gpt.adapter` surface (reference: lit_gpt/adapter.py) backed by lit_parrot_b200."""
from lit_parrot_b200.adapter import (  # noqa: F401
    GPT, Block, CausalSelfAttention, Config, adapter_filter, mark_only_adapter_as_trainable,
)
