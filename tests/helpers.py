"""Shared helpers for the tests (oracle = checker only)."""
import os

import numpy as np
import torch

from lit_parrot_b200 import Config

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TINY_NAMES = ["neox", "neox_seq", "falcon_mqa", "falcon_gqa", "llama_mha", "llama_gqa", "llama_condense"]


def load_tiny(name):
    z = np.load(os.path.join(GOLDEN, f"tiny_{name}.npz"), allow_pickle=False)
    kw = {k: eval(v) for k, v in zip(z["cfg_keys"].tolist(), z["cfg_vals"].tolist())}  # reprs of ints/floats/bools/strs
    return z, kw, Config(**kw)


def t(a, dtype=None):
    x = torch.from_numpy(np.asarray(a))
    return x if dtype is None else x.to(dtype)


def cosine(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-300))
