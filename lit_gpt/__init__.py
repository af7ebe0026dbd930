"""Drop-in import surface: `from lit_gpt import GPT, Config` resolves to the B200-native implementation."""
from lit_parrot_b200.config import Config
from lit_parrot_b200.model import GPT
from lit_parrot_b200.tokenizer import Tokenizer

__all__ = ["GPT", "Config", "Tokenizer"]
