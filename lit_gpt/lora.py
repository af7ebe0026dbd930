"""`lit_gpt.lora` surface (reference: lit_gpt/lora.py) backed by lit_parrot_b200 (merged-LoRA inference)."""
from lit_parrot_b200.lora import (  # noqa: F401
    GPT, Block, CausalSelfAttention, Config, GptNeoxMLP, LLaMAMLP, LoRALayer, LoRALinear, LoRAQKVLinear, lora_filter,
    mark_only_lora_as_trainable, merge_lora_weights,
)
