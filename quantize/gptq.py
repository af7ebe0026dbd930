"""`quantize.gptq` surface (reference: quantize/gptq.py): the int4 layer (205-264) and the GPTQ quantiser (267-609), backed by
lit_parrot_b200; `python quantize/gptq.py --checkpoint_dir ...` is the reference's CLI."""
from lit_parrot_b200.gptq import GPTQQuantizer, blockwise_quantization, get_sample_data, main  # noqa: F401
from lit_parrot_b200.quantize import ColBlockQuantizedLinear, rtn_int4_params  # noqa: F401

if __name__ == "__main__":
    from lit_parrot_b200.cli import CLI

    CLI(main)
