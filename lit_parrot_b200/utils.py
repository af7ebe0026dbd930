"""`quantization(mode)` — the reference's operator plug-in switch (lit_gpt/utils.py:26-83), same mode strings.

While the context is active ``torch.nn.Linear`` is replaced by a weight-only quantised class, so every
``nn.Linear(...)`` executed by ``GPT.__init__`` (including ``lm_head``) instantiates the plug-in; the original
class is restored on exit.
"""
from contextlib import contextmanager
from typing import Optional

import torch

from lit_parrot_b200.config import find_multiple  # noqa: F401  (re-exported like lit_gpt.utils.find_multiple)

_MODES = ("bnb.int8", "bnb.fp4", "bnb.fp4-dq", "bnb.nf4", "bnb.nf4-dq", "gptq.int4")


@contextmanager
def quantization(mode: Optional[str] = None, *, gptq_tile_cols: int = -1):
    """``gptq_tile_cols`` (extension): columns per scale/zero group of the GPTQ layer; the reference's context
    always builds per-row layers (tile_cols=-1, utils.py:72-74) although its quantiser supports groups."""
    if mode is None:
        yield
        return
    if mode not in _MODES:
        raise ValueError(f"Unknown quantization mode: {mode}")
    if "fp4" in mode or mode.endswith("-dq"):
        # accepted by the reference (bitsandbytes arithmetic that is not restated here, DESIGN.md "out of scope"): refuse at mode
        # selection, not in the middle of GPT construction
        raise NotImplementedError(f"quantization mode {mode!r}: bitsandbytes FP4 and double quantisation (-dq) have no B200 kernel; "
                                  "use 'bnb.nf4', 'bnb.int8' or 'gptq.int4'")
    from lit_parrot_b200 import quantize as q

    if mode == "bnb.int8":
        quantized_linear_cls = q.InferenceLinear8bitLt
    elif mode.startswith("bnb."):
        quant_type = "fp4" if "fp4" in mode else "nf4"
        dq = mode.endswith("-dq")

        class QuantizedLinear(q.Linear4bit):
            def __init__(self, *args, **kwargs):
                super().__init__(*args, quant_type=quant_type, compress_statistics=dq, **kwargs)

        quantized_linear_cls = QuantizedLinear
    else:

        class QuantizedLinear(q.ColBlockQuantizedLinear):
            def __init__(self, *args, **kwargs):
                super().__init__(*args, bits=4, tile_cols=gptq_tile_cols, **kwargs)

        quantized_linear_cls = QuantizedLinear

    torch_linear_cls = torch.nn.Linear
    torch.nn.Linear = quantized_linear_cls
    try:
        yield
    finally:
        torch.nn.Linear = torch_linear_cls
