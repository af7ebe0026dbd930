"""`quantize.gptq.ColBlockQuantizedLinear` surface (reference: quantize/gptq.py:205-264)."""
from lit_parrot_b200.quantize import ColBlockQuantizedLinear, rtn_int4_params  # noqa: F401
