"""`generate.base.generate` surface (reference: generate/base.py:92-159) backed by lit_parrot_b200."""
from lit_parrot_b200.generate import generate  # noqa: F401
