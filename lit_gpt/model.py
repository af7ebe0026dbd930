"""`lit_gpt.model` surface (reference: lit_gpt/model.py) backed by lit_parrot_b200."""
from lit_parrot_b200.model import (  # noqa: F401
    GPT, Block, CausalSelfAttention, GptNeoxMLP, KVCache, LLaMAMLP, RoPECache, build_rope_cache,
)
