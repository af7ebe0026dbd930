"""`lit_gpt.tokenizer.Tokenizer` (reference: lit_gpt/tokenizer.py)."""
from lit_parrot_b200.tokenizer import Tokenizer  # noqa: F401
