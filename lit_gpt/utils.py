"""`lit_gpt.utils` surface: only the hot-path members (`quantization`, `find_multiple`)."""
from lit_parrot_b200.utils import find_multiple, quantization  # noqa: F401
