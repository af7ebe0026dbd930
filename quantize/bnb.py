"""`quantize.bnb` surface (reference: quantize/bnb.py:18-75): NF4 / int8 weight-only layers."""
from lit_parrot_b200.quantize import InferenceLinear8bitLt, Linear4bit  # noqa: F401
