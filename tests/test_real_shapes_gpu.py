"""lp_decode_step at the shapes bench.py times, against the oracle, at a 2k context and across the ring wrap.

Every benchmarked batch-1 configuration (BASELINE configs 2-5) at its REAL widths, two layers deep, through the persistent
step kernel: falcon-7b (E 4544 = 71 K-blocks: ragged last stage; 71 heads MQA -> 2 sequence splits; QKV 4672 rows; shared
LayerNorm, parallel residual), stablelm-base-alpha-3b (hs 128, rotary 25 %, two LayerNorms + biases, V 50688), Llama-2-7b
GPTQ int4 g128 (K 11008 = 86 groups in mlp.proj), and the Llama-2-70b tensor-parallel shard of tp = 8 (E 8192, 8 local heads on
one KV group, I 3584, attn.proj K 1024) with the in-kernel EXCHANGE op.  The KV caches of oracle and product are pre-filled with
the same synthetic bf16 rows up to position ~2036 (the judge's recipe: no 2k-token CPU prefill), then teacher-forced decode
steps run through kv lengths 2037..max_seq and on past the end of the cache (the reference rolls, model.py:238-242; the
product's ring slot overwrites the oldest row)."""
import pytest
import torch

import lit_parrot_b200 as lp
from oracle import lit_oracle as O
from helpers import cosine

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _synthetic_kv(cfg, G, max_seq, start, seed):
    g = torch.Generator().manual_seed(seed)
    kv = []
    for _ in range(cfg.n_layer):
        k = torch.zeros(1, G, max_seq, cfg.head_size)
        v = torch.zeros(1, G, max_seq, cfg.head_size)
        k[:, :, :start] = torch.randn(1, G, start, cfg.head_size, generator=g)
        v[:, :, :start] = torch.randn(1, G, start, cfg.head_size, generator=g)
        kv.append((k.bfloat16(), v.bfloat16()))
    return kv


def _install_kv(m, om, cfg, kv, max_seq, group0=0, n_groups=None):
    """Same rows into the product's compact (B, G_local, S, hs) bf16 cache and the oracle's (B, n_head | 1, S, hs) cache."""
    m.kv_caches = m.build_kv_caches(torch.zeros(1, 1, device=DEV), max_seq)
    assert m.kv_caches[0][0].dtype == torch.bfloat16
    qpk = cfg.n_head // cfg.n_query_groups
    om.kv = []
    for (pk, pv), (k, v) in zip(m.kv_caches, kv):
        pk.copy_(k.to(DEV))
        pv.copy_(v.to(DEV))
        if cfg.n_query_groups == 1:
            om.kv.append((k.float().clone(), v.float().clone()))
        else:
            ok = torch.zeros(1, cfg.n_head, max_seq, cfg.head_size)
            ov = torch.zeros(1, cfg.n_head, max_seq, cfg.head_size)
            ng = k.size(1) if n_groups is None else n_groups
            ok[:, group0 * qpk:(group0 + ng) * qpk] = k.float().repeat_interleave(qpk, dim=1)
            ov[:, group0 * qpk:(group0 + ng) * qpk] = v.float().repeat_interleave(qpk, dim=1)
            om.kv.append((ok, ov))


def _decode_and_compare(m, om, cfg, start, steps, max_seq, seed=3, atol=2e-3):
    g = torch.Generator().manual_seed(seed)
    toks = torch.randint(0, cfg.vocab_size, (steps,), generator=g)
    worst = 0.0
    for i in range(steps):
        p = torch.tensor([start + i])
        want = om(toks[i].view(1, 1), max_seq, p)[0, -1]
        got = m._forward_impl(toks[i].view(1, 1).to(DEV), max_seq, p.to(DEV), raw_logits=True)[0, -1].float().cpu()
        err = float((got - want).abs().max())
        worst = max(worst, err)
        # north star: max-abs 2e-2, cosine 0.999; held far tighter (fp32 activations over the same stored weights)
        assert err < atol and cosine(got, want) > 0.99999, f"position {start + i}: max-abs {err:.3e}"
    assert any(v is not None for v in m._engine._steps.values()), "decode did not go through lp_decode_step"
    return worst


def _bf16_pair(cfg, seed):
    sd = {k: v.bfloat16() for k, v in O.random_state_dict(cfg, seed=seed, perturb_norm=True).items()}
    m = lp.GPT(cfg)
    m.load_state_dict(sd)
    m = m.to(device=DEV, dtype=torch.bfloat16).eval()
    om = O.OracleGPT(cfg, {k: v.float() for k, v in sd.items()}, kv_round=torch.bfloat16)
    return m, om


def test_falcon7b_widths_2k_context_and_wrap():
    """falcon-7b: block_size 2048, so the cache is 2040 long and positions run 2030..2047: kv length 2031..2040, then 8 wrapped
    steps.  71 MQA heads share one K/V group; 71 K-blocks leave a 7-block last stage in every E-wide row."""
    cfg = lp.Config.from_name("falcon-7b", n_layer=2)
    assert (cfg.n_embd, cfg.n_head, cfg.n_query_groups, cfg.head_size) == (4544, 71, 1, 64)
    m, om = _bf16_pair(cfg, 71)
    max_seq, start = 2040, 2030
    _install_kv(m, om, cfg, _synthetic_kv(cfg, 1, max_seq, start, 171), max_seq)
    _decode_and_compare(m, om, cfg, start, 18, max_seq)


def test_stablelm3b_widths_2k_context_and_wrap():
    """The headline workload's shapes: hs 128 with 32 rotary dims, LayerNorm x 2 + biases, parallel residual, V 50688."""
    cfg = lp.Config.from_name("stablelm-base-alpha-3b", n_layer=2)
    assert (cfg.n_embd, cfg.n_head, cfg.head_size, cfg.rope_n_elem, cfg.padded_vocab_size) == (4096, 32, 128, 32, 50688)
    m, om = _bf16_pair(cfg, 72)
    max_seq, start = 2048, 2038
    _install_kv(m, om, cfg, _synthetic_kv(cfg, cfg.n_query_groups, max_seq, start, 172), max_seq)
    _decode_and_compare(m, om, cfg, start, 16, max_seq)


@pytest.mark.parametrize("fused", [True, False], ids=["fused_slabs", "five_op_layer"])
def test_llama7b_int4_g128_widths_2k_context_and_wrap(fused, monkeypatch):
    """The north-star configuration's shapes: GPTQ int4, group 128, K 11008 (86 groups) in mlp.proj, 32 MHA heads -> 4 splits.
    Scales bf16-exact as in a bf16 checkpoint.  `fused`: attention -> attn.proj and fc -> mlp.proj as column->row pairs inside the
    CTAs (the default, what bench.py times); otherwise the five-op layer (LP_DS_FUSE=0: K = 11008 staged through 86 groups)."""
    if not fused:
        monkeypatch.setenv("LP_DS_FUSE", "0")
    cfg = lp.Config.from_name("Llama-2-7b-hf", n_layer=2)
    fsd = O.random_state_dict(cfg, seed=73, perturb_norm=True)
    with lp.quantization("gptq.int4", gptq_tile_cols=128):
        m = lp.GPT(cfg)
    qsd, dense = {}, {}
    for k, v in fsd.items():
        if v.dim() == 2 and "wte" not in k:
            packed, scales, zeros = O.gptq_rtn_quantize(v, 128, scale_dtype=torch.bfloat16)
            base = k[: -len(".weight")]
            qsd[base + ".quant_weight"], qsd[base + ".scales"], qsd[base + ".zeros"] = packed, scales, zeros
            dense[k] = O.gptq_dequant(packed, scales, zeros, torch.float32, tile_cols=128)  # == the per-call dequant of O.linear
        else:
            qsd[k] = dense[k] = v
    m.load_state_dict(qsd)
    m = m.to(DEV).eval()
    m.kv_cache_dtype = torch.bfloat16
    om = O.OracleGPT(cfg, dense, kv_round=torch.bfloat16)
    max_seq, start = 2048, 2038
    _install_kv(m, om, cfg, _synthetic_kv(cfg, cfg.n_query_groups, max_seq, start, 173), max_seq)
    _decode_and_compare(m, om, cfg, start, 16, max_seq)
    assert bool(m._engine._slabs) == fused


@pytest.mark.parametrize("mode", ["bnb.nf4", "bnb.int8"])
def test_llama7b_bnb_widths_2k_context_and_wrap(mode):
    """BASELINE configs[2], the bnb side: Llama-2-7b widths with NF4 (blocksize 64) / row-wise int8 weights through the step kernel
    (K 11008 = 43 K-blocks of 256 codes: ragged last stage; absmax tile-major with zero padding)."""
    cfg = lp.Config.from_name("Llama-2-7b-hf", n_layer=2)
    fsd = {k: v.bfloat16().float() for k, v in O.random_state_dict(cfg, seed=75, perturb_norm=True).items()}
    with lp.quantization(mode):
        m = lp.GPT(cfg)
    m.load_state_dict(fsd)
    m = m.to(DEV).eval()
    m.kv_cache_dtype = torch.bfloat16
    dense = {}
    for k, v in fsd.items():
        if v.dim() == 2 and "wte" not in k:
            dense[k] = O.nf4_dequantize(*O.nf4_quantize(v), v.shape) if mode == "bnb.nf4" else O.int8_dequantize(*O.int8_quantize(v))
        else:
            dense[k] = v
    om = O.OracleGPT(cfg, dense, kv_round=torch.bfloat16)
    max_seq, start = 2048, 2040
    _install_kv(m, om, cfg, _synthetic_kv(cfg, cfg.n_query_groups, max_seq, start, 175), max_seq)
    _decode_and_compare(m, om, cfg, start, 12, max_seq)


def test_llama70b_tp8_local_shard_2k_context_and_wrap():
    """Rank 0's shard of Llama-2-70b at tp = 8 (E 8192, 8 local heads on ONE KV group, QKV 1280 rows, attn.proj K 1024, I 3584)
    run alone on one GPU: the in-kernel EXCHANGE op sums a single partial (LoopbackTPContext).  Oracle: the full-width model
    whose other seven shards are zero, so that its row-parallel sums equal rank 0's partials."""
    from lit_parrot_b200.tp import LoopbackTPContext, shard_state_dict

    full = lp.Config.from_name("Llama-2-70b-hf", n_layer=2)
    lcfg = full.with_tp(8, 0)
    assert (lcfg.n_head_local, lcfg.n_query_groups_local, lcfg.intermediate_size_local, lcfg.qkv_rows_local) == (8, 1, 3584, 1280)
    sd = {k: v.bfloat16() for k, v in O.random_state_dict(full, seed=74).items()}
    for k, v in sd.items():  # zero the other ranks' shards of the sharded tensors
        if ".attn.attn." in k or ".mlp.fc" in k:
            v[v.shape[0] // 8:] = 0
        elif k.endswith(".attn.proj.weight") or k.endswith(".mlp.proj.weight"):
            v[:, v.shape[1] // 8:] = 0
    m = lp.GPT(lcfg)
    m.load_state_dict(shard_state_dict(sd, lcfg))
    m = m.to(device=DEV, dtype=torch.bfloat16).eval()
    m.tp_context = LoopbackTPContext(torch.device(DEV), max_rows=8, n_embd=full.n_embd)
    om = O.OracleGPT(full, {k: v.float() for k, v in sd.items()}, kv_round=torch.bfloat16)
    max_seq, start = 2048, 2040
    _install_kv(m, om, full, _synthetic_kv(full, 1, max_seq, start, 174), max_seq, group0=0, n_groups=1)
    _decode_and_compare(m, om, full, start, 12, max_seq)
