"""`lit_gpt.rmsnorm` surface backed by lit_parrot_b200."""
from lit_parrot_b200.model import RMSNorm  # noqa: F401
