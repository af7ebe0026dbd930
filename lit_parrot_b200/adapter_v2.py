"""LLaMA-Adapter v2 inference — drop-in surface of the reference's ``lit_gpt/adapter_v2.py``.

v2 gives every ``nn.Linear`` of the model a per-output ``adapter_bias`` and ``adapter_scale`` and evaluates
``adapter_scale * (linear(x) + adapter_bias)`` (adapter_v2.py:34-35).  Here the two vectors are parameters of the (compute-free)
linear containers under the reference's names, and the arithmetic is the ``out_bias`` / ``out_scale`` output affine of
``lp_weight``: applied inside the epilogues of the GEMV / GEMM kernels, in the reference's order and with its bf16 rounding
points, before the activation / residual — no extra launch, no extra pass over the activations.
"""
from typing import Any

import torch

from lit_parrot_b200.adapter import GPT

_V2_KEYS = ("adapter_wte", "gating_factor",  # adapter v1
            "adapter_scale", "adapter_bias",  # v2: output affine of every linear layer
            "norm_1", "norm_2", "ln_f")  # v2: the norms are trainable too


def adapter_filter(key: str, value: Any) -> bool:
    """Which state-dict entries an adapter-v2 checkpoint holds (adapter_v2.py:12-25)."""
    return any(s in key for s in _V2_KEYS)


def mark_only_adapter_v2_as_trainable(model: GPT) -> None:
    for name, param in model.named_parameters():
        param.requires_grad = adapter_filter(name, param)


def adapter_v2_linear_with_bias_and_scale(layer: torch.nn.Module) -> torch.nn.Module:
    """Zero bias / unit scale in the weight's dtype, frozen, under the reference's parameter names (adapter_v2.py:38-47)."""
    w = layer.weight
    layer.adapter_bias = torch.nn.Parameter(torch.zeros(w.shape[0], dtype=w.dtype, device=w.device), requires_grad=False)
    layer.adapter_scale = torch.nn.Parameter(torch.ones(w.shape[0], dtype=w.dtype, device=w.device), requires_grad=False)
    return layer


def add_adapter_v2_parameters_to_linear_layers(model: torch.nn.Module) -> None:
    for module in model.modules():
        if isinstance(module, torch.nn.Linear):
            if hasattr(module, "lp_pack"):
                raise NotImplementedError("adapter v2 on weight-only quantised layers is not supported")
            adapter_v2_linear_with_bias_and_scale(module)
    if hasattr(model, "invalidate"):
        model.invalidate()  # packed weight records of an existing engine do not know the new vectors
