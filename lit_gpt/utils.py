"""`lit_gpt.utils` surface: the hot-path members (`quantization`, `find_multiple`) and checkpoint interop
(`lazy_load`, `check_valid_checkpoint_dir`)."""
from lit_parrot_b200.checkpoint import check_valid_checkpoint_dir, lazy_load  # noqa: F401
from lit_parrot_b200.utils import find_multiple, quantization  # noqa: F401
