"""Host-side logic that needs no GPU: module tree / state-dict contract, plug-in switch, error behaviour,
generate() driving a foreign (CPU) model with the reference's semantics."""
import os
from unittest import mock

import numpy as np
import pytest
import torch

import lit_parrot_b200 as lp
from oracle import lit_oracle as O
from helpers import TINY_NAMES, load_tiny, t


@pytest.mark.parametrize("name", TINY_NAMES + ["pythia-70m", "falcon-7b"])
def test_state_dict_contract(name):
    if name in TINY_NAMES:
        _, _, cfg = load_tiny(name)
    else:
        cfg = lp.Config.from_name(name, n_layer=1)
    with torch.device("meta"):
        m = lp.GPT(cfg)
    got = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert got == O.state_dict_shapes(cfg)


def test_drop_in_import_surface():
    import generate.base
    import lit_gpt
    import lit_gpt.model
    import lit_gpt.utils
    import quantize.bnb
    import quantize.gptq

    assert lit_gpt.GPT is lp.GPT and lit_gpt.Config is lp.Config
    assert generate.base.generate is lp.generate
    assert lit_gpt.utils.quantization is lp.quantization
    assert lit_gpt.utils.find_multiple(50254, 512) == 50688
    assert hasattr(lit_gpt.model, "build_rope_cache") and hasattr(quantize.gptq, "ColBlockQuantizedLinear")
    assert hasattr(quantize.bnb, "Linear4bit") and hasattr(quantize.bnb, "InferenceLinear8bitLt")


def test_quantization_plugin_switch(golden_dir):
    """lit_gpt/utils.py:26-83: torch.nn.Linear is swapped during construction (lm_head included) and restored after;
    the GPTQ state-dict keys/shapes equal the reference's."""
    _, kw, cfg = load_tiny("llama_mha")
    linear = torch.nn.Linear
    with lp.quantization("gptq.int4"):
        assert torch.nn.Linear is not linear
        m = lp.GPT(cfg)
    assert torch.nn.Linear is linear
    z = np.load(f"{golden_dir}/gptq_statedict_keys.npz")
    ref = dict(zip(z["keys"].tolist(), z["shapes"].tolist()))
    got = {k: repr(tuple(v.shape)) for k, v in m.state_dict().items()}
    assert got == ref
    qw = m.lm_head.quant_weight
    assert qw.dtype == torch.uint8 and qw.stride() == (1, qw.shape[0])  # reference storage order (gptq.py:216-222)
    with pytest.raises(ValueError, match="Unknown quantization mode"):
        with lp.quantization("nope"):
            pass
    assert torch.nn.Linear is linear
    with lp.quantization(None):
        assert torch.nn.Linear is linear
    for mode, cls in (("bnb.nf4", "Linear4bit"), ("bnb.int8", "InferenceLinear8bitLt")):
        with lp.quantization(mode):
            m = lp.GPT(lp.Config(block_size=16, vocab_size=64, padding_multiple=64, n_layer=1, n_head=2, n_embd=64))
        assert cls in [c.__name__ for c in type(m.lm_head).__mro__]
        assert torch.nn.Linear is linear


def test_gptq_module_matches_reference_fixture(golden_dir):
    """ColBlockQuantizedLinear: RTN quantiser, nibble order and get_weight equal the reference fixture (CPU, torch ops)."""
    from lit_parrot_b200.quantize import ColBlockQuantizedLinear

    for tag in ("g128", "perrow"):
        z = np.load(f"{golden_dir}/gptq_{tag}.npz")
        w = t(z["w"])
        lin = ColBlockQuantizedLinear(w.shape[1], w.shape[0], True, bits=4, tile_cols=int(z["tile_cols"]))
        lin.quantize_rtn_(w)
        assert torch.equal(lin.quant_weight, t(z["quant_weight"]))
        assert torch.equal(lin.scales, t(z["scales"])) and torch.equal(lin.zeros, t(z["zeros"]))
        assert torch.equal(lin.get_weight(torch.float32), t(z["dequant"]))


def test_cpu_is_refused_loudly():
    cfg = lp.Config(block_size=16, vocab_size=32, padding_multiple=32, n_layer=1, n_head=2, n_embd=16)
    m = lp.GPT(cfg)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 4, dtype=torch.long))
    with pytest.raises(RuntimeError, match="CUDA"):
        lp.generate(m, torch.zeros(4, dtype=torch.int32), 8)
    # the reference's assertions come first (model.py:72-77)
    with pytest.raises(AssertionError, match="block size is only"):
        m(torch.zeros(1, 17, dtype=torch.long))
    with pytest.raises(AssertionError, match="max seq length is only"):
        m(torch.zeros(1, 8, dtype=torch.long), 4, torch.arange(8))


@pytest.mark.parametrize("max_seq_length", (10, 20 + 5))
def test_generate_foreign_model_reference_semantics(max_seq_length):
    """reference tests/test_generate.py:14-42, with the oracle as the (foreign, CPU) model: output == prompt ++ samples."""
    T = 5
    input_idx = torch.randint(10, size=(T,))
    cfg = lp.Config(block_size=128, vocab_size=16, n_layer=1, n_head=4, n_embd=8)
    model = O.OracleGPT(cfg, O.random_state_dict(cfg, seed=3))
    results = []
    original = torch.multinomial

    def multinomial(*args, **kwargs):
        out = original(*args, **kwargs)
        results.append(out)
        return out

    with mock.patch("torch.multinomial", multinomial):
        out = lp.generate(model, input_idx, T + 20, max_seq_length=max_seq_length, top_k=4)
    assert out.size(0) == T + 20
    torch.testing.assert_close(out, torch.cat((input_idx, torch.hstack(results))))


def test_generate_foreign_eos_cut():
    cfg = lp.Config(block_size=64, vocab_size=16, n_layer=1, n_head=2, n_embd=16)
    model = O.OracleGPT(cfg, O.random_state_dict(cfg, seed=5))
    prompt = torch.tensor([1, 2, 3], dtype=torch.int32)
    full = lp.generate(model, prompt, 20, top_k=1)
    model.reset_cache()
    eos = int(full[7])
    first = int((full[3:] == eos).nonzero()[0]) + 3
    cut = lp.generate(model, prompt, 20, top_k=1, eos_id=eos)
    assert torch.equal(cut, full[:first])  # idx[:input_pos] excludes the EOS itself (generate/base.py:156-157)


def test_step_kernel_register_budget():
    """decode_step_kernel is ONE function holding every op kind of the step (linear formats, attention, slab, exchange); its register
    allocation is global, and an innocent-looking change in a cold path once pushed spills into the hot loops (stablelm-3b 690 ->
    640 tok/s with 626 instead of 264 bytes of spill stores).  Keep the spill volume of the shipped source under watch."""
    import re
    import subprocess

    from lit_parrot_b200 import build as b

    src = os.path.join(b.CSRC, "decode_step.cu")
    cmd = [b._nvcc(), *b.NVCC_FLAGS, "-Xptxas=-v", "-c", src, "-o", os.devnull]
    out = subprocess.run(cmd, capture_output=True, text=True).stderr
    blocks = re.findall(r"Function properties for (\S*decode_step_kernel\S*)\s+(\d+) bytes stack frame, (\d+) bytes spill stores", out)
    assert len(blocks) == 2, out[-2000:]
    for name, _stack, stores in blocks:
        assert int(stores) <= 400, f"{name}: {stores} bytes of spill stores (budget 400; 264 when this test was written)"


def test_swap_gemm_code_footprint():
    """gemm_tc_swap_kernel runs its tile epilogue ONCE per CTA on single-tile shapes; as one unrolled body holding every variant it was
    17.2 k SASS instructions and a launch spent as long fetching them (20 us, all SMs walking the same cold lines in lock step) as
    streaming its weights (18 us).  The pass-structured epilogue is 5.2 k; keep the whole kernel under 8 k."""
    import re
    import subprocess

    from lit_parrot_b200 import build as b

    obj = os.path.join(b.HERE, "build", "gemm_tc.o")
    src = os.path.join(b.CSRC, "gemm_tc.cu")
    if not os.path.exists(obj) or os.path.getmtime(obj) < os.path.getmtime(src):
        subprocess.run([b._nvcc(), *b.NVCC_FLAGS, "-c", src, "-o", obj], check=True, capture_output=True)
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    counts, name = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
        elif name and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", line):
            counts[name] = counts.get(name, 0) + 1
    swap = {k: v for k, v in counts.items() if "gemm_tc_swap_kernel" in k}
    assert len(swap) == 4, sorted(counts)  # <affine, stream-K> variants
    for k, v in swap.items():
        assert v < 8000, f"{k}: {v} SASS instructions (5.2 k when this test was written)"
