"""`lit_gpt.adapter_v2` surface (reference: lit_gpt/adapter_v2.py) backed by lit_parrot_b200."""
from lit_parrot_b200.adapter import GPT  # noqa: F401
from lit_parrot_b200.adapter_v2 import (  # noqa: F401
    adapter_filter, adapter_v2_linear_with_bias_and_scale, add_adapter_v2_parameters_to_linear_layers,
    mark_only_adapter_v2_as_trainable,
)
