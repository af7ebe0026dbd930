"""`lit_gpt.config` surface (reference: lit_gpt/config.py) backed by lit_parrot_b200."""
from lit_parrot_b200.config import Config, configs, name_to_config  # noqa: F401
