"""lit_parrot_b200 — B200-native (sm_100a) implementation of the Lit-GPT inference hot path.

Drop-in surface: ``GPT`` / ``Config`` (lit_gpt.model / lit_gpt.config), ``generate`` (generate/base.py),
``quantization`` (lit_gpt.utils).  All device arithmetic is done by liblitparrot_b200.so through the C ABI in
include/lp_abi.h; importing this package does not need a GPU, running a model does.
"""
from lit_parrot_b200.config import Config, name_to_config  # noqa: F401
from lit_parrot_b200.model import GPT  # noqa: F401
from lit_parrot_b200.generate import generate, sample  # noqa: F401
from lit_parrot_b200.utils import quantization  # noqa: F401
from lit_parrot_b200.checkpoint import check_valid_checkpoint_dir, lazy_load, load_checkpoint, save_checkpoint  # noqa: F401

__all__ = ["GPT", "Config", "generate", "sample", "quantization", "name_to_config", "lazy_load", "check_valid_checkpoint_dir",
           "load_checkpoint", "save_checkpoint"]
