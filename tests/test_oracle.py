"""The CPU oracle (oracle/lit_oracle.py) against the golden vectors produced by the unmodified reference."""
import numpy as np
import pytest
import torch

from oracle import lit_oracle as O
from helpers import TINY_NAMES, load_tiny, t


@pytest.mark.parametrize("name", TINY_NAMES)
def test_oracle_matches_reference_golden(name):
    z, kw, cfg = load_tiny(name)
    sd = O.random_state_dict(cfg, seed=int(z["seed"]), perturb_norm=True)
    m = O.OracleGPT(cfg, sd)
    idx = t(z["idx"])
    assert torch.equal(m(idx), t(z["ref_full"]))
    max_seq = int(z["max_seq"])
    pos = torch.arange(idx.shape[1])
    assert torch.equal(m(idx, max_seq, pos), t(z["ref_prefill"]))
    for s, tok in enumerate(t(z["forced"])):
        pos = pos[-1:] + 1
        assert torch.equal(m(tok, max_seq, pos), t(z["ref_steps"][s]))
    m.reset_cache()
    prompt = t(z["prompt"])
    assert torch.equal(O.generate(m, prompt, 30, 30, top_k=1, argmax_ties=True), t(z["ref_gen"]))
    m.reset_cache()
    assert torch.equal(O.generate(m, prompt, 30, 12, top_k=1, argmax_ties=True), t(z["ref_gen_overflow"]))


@pytest.mark.parametrize("tag", ["g128", "perrow"])
def test_oracle_gptq_golden(tag, golden_dir):
    z = np.load(f"{golden_dir}/gptq_{tag}.npz")
    tile = int(z["tile_cols"])
    packed, scales, zeros = O.gptq_rtn_quantize(t(z["w"]), tile)
    assert torch.equal(packed, t(z["quant_weight"]))
    assert torch.equal(scales, t(z["scales"])) and torch.equal(zeros, t(z["zeros"]))
    assert torch.equal(O.gptq_dequant(packed, scales, zeros), t(z["dequant"]))
    assert torch.equal(O.gptq_dequant(packed, scales, zeros, dtype=torch.bfloat16).float(), t(z["dequant_bf16_as_f32"]))
    sd = {"l.quant_weight": packed, "l.scales": scales, "l.zeros": zeros, "l.bias": t(z["bias"])}
    assert torch.equal(O.linear(t(z["x"]), sd, "l"), t(z["y"]))


def test_oracle_pythia70m_greedy(golden_dir):
    """BASELINE config 1: pythia-70m random init fp32, greedy 16 -> 128 tokens."""
    from lit_parrot_b200 import Config

    z = np.load(f"{golden_dir}/pythia70m_greedy.npz")
    cfg = Config.from_name("pythia-70m")
    sd = O.random_state_dict(cfg, seed=int(z["seed"]))
    m = O.OracleGPT(cfg, sd)
    logits = []
    toks = O.generate(m, t(z["prompt"]), 128, 128, top_k=1, argmax_ties=True, logits_out=logits)
    assert torch.equal(toks, t(z["tokens"]))
    lg = torch.stack(logits)
    probe = z["probe_steps"].tolist()
    torch.testing.assert_close(lg[probe], t(z["probe_logits"]), rtol=0, atol=1e-5)


def test_nf4_and_int8_roundtrip():
    """NF4 / int8 restatements (parity unpinned): codebook is symmetric-ish, dequant error is bounded by the grid."""
    g = torch.Generator().manual_seed(0)
    w = torch.randn(32, 128, generator=g) * 0.02
    packed, absmax = O.nf4_quantize(w)
    assert packed.dtype == torch.uint8 and packed.numel() == w.numel() // 2 and absmax.numel() == w.numel() // 64
    wd = O.nf4_dequantize(packed, absmax, w.shape)
    # largest gap between neighbouring NF4 codes is < 0.31, so |err| <= 0.155 * absmax
    err = (wd - w).abs().reshape(-1, 64).amax(1)
    assert torch.all(err <= 0.16 * absmax + 1e-8)
    # first element of each pair sits in the high nibble
    first = (w.reshape(-1)[0] / absmax[0] - O.NF4_CODE).abs().argmin()
    assert int(packed[0] >> 4) == int(first)
    cb, scb = O.int8_quantize(w)
    assert cb.dtype == torch.int8 and torch.all(cb.abs().amax(1) == 127)
    assert (O.int8_dequantize(cb, scb) - w).abs().max() <= scb.max() / 127 * 0.5 + 1e-8
