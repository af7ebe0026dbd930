"""The whole-step persistent decode kernel (lp_decode_step, csrc/decode_step.cu) against the oracle and against the per-op
path of the same library: Llama-style (RMSNorm, SwiGLU, sequential residual) and NeoX-style (LayerNorm + bias, GELU, parallel
residual, partial rotary) blocks, bf16 and GPTQ-int4 weights, head sizes 128 and 64, contexts that cross key-tile
boundaries and wrap the cache."""
import pytest
import torch

import lit_parrot_b200 as lp
from oracle import lit_oracle as O
from helpers import cosine

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

LLAMA = dict(block_size=256, vocab_size=320, padding_multiple=64, n_layer=3, n_head=2, n_embd=256, rotary_percentage=1.0,
             parallel_residual=False, bias=False, _norm_class="RMSNorm", _mlp_class="LLaMAMLP", intermediate_size=512)
NEOX = dict(block_size=256, vocab_size=320, padding_multiple=64, n_layer=3, n_head=4, n_embd=256, rotary_percentage=0.25,
            parallel_residual=True, bias=True, _norm_class="LayerNorm", _mlp_class="GptNeoxMLP")
NEOX_SHARED = dict(NEOX, shared_attention_norm=True, n_head=2)
LLAMA_GQA = dict(LLAMA, n_head=4, n_query_groups=2)                                    # hs 64, two q heads per KV group
LLAMA_GQA128 = dict(LLAMA, n_embd=512, n_head=4, n_query_groups=2, intermediate_size=768)  # hs 128
FALCON_MQA = dict(NEOX, n_head=4, n_query_groups=1, rotary_percentage=1.0, shared_attention_norm=True, bias=False)  # falcon-7b style


def bf16_model(kw, seed):
    cfg = lp.Config(**kw)
    sd = {k: v.bfloat16() for k, v in O.random_state_dict(cfg, seed=seed, perturb_norm=True).items()}
    m = lp.GPT(cfg)
    m.load_state_dict(sd)
    m = m.to(device=DEV, dtype=torch.bfloat16).eval()
    om = O.OracleGPT(cfg, {k: v.float() for k, v in sd.items()}, kv_round=torch.bfloat16)
    return cfg, m, om


def int4_model(kw, seed, tile=128, scale_dtype=None):
    cfg = lp.Config(**kw)
    fsd = O.random_state_dict(cfg, seed=seed, perturb_norm=True)
    with lp.quantization("gptq.int4", gptq_tile_cols=tile):
        m = lp.GPT(cfg)
    qsd = {}
    for k, v in fsd.items():
        if v.dim() == 2 and "wte" not in k:
            packed, scales, zeros = O.gptq_rtn_quantize(v, tile, scale_dtype=scale_dtype)
            base = k[: -len(".weight")]
            qsd[base + ".quant_weight"], qsd[base + ".scales"], qsd[base + ".zeros"] = packed, scales, zeros
        else:
            qsd[k] = v
    m.load_state_dict(qsd)
    m = m.to(DEV).eval()
    m.kv_cache_dtype = torch.bfloat16
    return cfg, m, O.OracleGPT(cfg, qsd, kv_round=torch.bfloat16)


def step_kernel_used(m) -> bool:
    return any(v is not None for v in m._engine._steps.values())


def teacher_forced(m, om, cfg, prompt_len, steps, max_seq, seed=5, prompt_atol=5e-4):
    g = torch.Generator().manual_seed(seed)
    toks = torch.randint(0, cfg.padded_vocab_size, (prompt_len + steps,), generator=g)
    pos = torch.arange(prompt_len)
    want = om(toks[:prompt_len].view(1, -1), max_seq, pos)[0, -1]
    got = m._forward_impl(toks[:prompt_len].view(1, -1).to(DEV), max_seq, pos.to(DEV), raw_logits=True)[0, -1].float().cpu()
    # the 60-row prompt runs on the swap-AB GEMM with two bf16 terms per activation (2^-17 relative per product) and atomically
    # accumulated split-K partials (order varies from run to run): observed 1.2e-4 .. 2.0e-4; north star 2e-2
    torch.testing.assert_close(got, want, rtol=0, atol=prompt_atol)
    worst = 0.0
    for i in range(prompt_len, prompt_len + steps):
        p = torch.tensor([i])
        want = om(toks[i].view(1, 1), max_seq, p)[0, -1]
        got = m._forward_impl(toks[i].view(1, 1).to(DEV), max_seq, p.to(DEV), raw_logits=True)[0, -1].float().cpu()
        err = float((got - want).abs().max())
        worst = max(worst, err)
        # bf16 cache: a k / v element that sits on a bf16 rounding boundary may round the other way than in the oracle
        # (1 bf16 ulp of one element), hence 1e-3 rather than the 2e-5 of the fp32-cache tests; north star: 2e-2 / 0.999
        assert err < 1e-3 and cosine(got, want) > 0.99999, f"step {i}: max-abs {err:.3e}"
    return worst


@pytest.mark.parametrize("kw,seed", [(LLAMA, 51), (NEOX, 52), (NEOX_SHARED, 53), (LLAMA_GQA, 58), (LLAMA_GQA128, 59), (FALCON_MQA, 60)],
                         ids=["llama_hs128", "neox_hs64", "neox_shared_hs128", "llama_gqa_hs64", "llama_gqa_hs128", "falcon_mqa_hs64"])
def test_step_kernel_logits_bf16(kw, seed):
    """Teacher-forced decode from position 60 to 150: key tiles fill up, a second (third) tile and sequence split appear."""
    cfg, m, om = bf16_model(kw, seed)
    teacher_forced(m, om, cfg, prompt_len=60, steps=90, max_seq=256)
    assert step_kernel_used(m), "decode step did not go through lp_decode_step"


def test_step_kernel_logits_int4():
    cfg, m, om = int4_model(LLAMA, 54)
    teacher_forced(m, om, cfg, prompt_len=60, steps=80, max_seq=256)
    assert step_kernel_used(m)


@pytest.mark.parametrize("kw,seed", [(NEOX, 61), (NEOX_SHARED, 62)], ids=["neox", "neox_shared_norm"])
def test_step_kernel_logits_int4_layernorm_parallel_residual(kw, seed):
    """GPTQ-int4 weights behind LayerNorm in a parallel-residual block: QKV and FC stage the same row back to back (the second
    from the raw row + statistics the first leaves in shared memory), through the int8-digit activation path."""
    cfg, m, om = int4_model(kw, seed)
    teacher_forced(m, om, cfg, prompt_len=40, steps=40, max_seq=256)
    assert step_kernel_used(m)


@pytest.mark.parametrize("which", ["bf16", "int4", "gqa", "mqa"])
def test_step_kernel_greedy_tokens_and_sliding_window(which):
    """generate(): 100 greedy tokens identical to the oracle, then the overflow case (max_seq_length 80 < 140 tokens: the ring
    slot replaces the reference's roll, model.py:238-242)."""
    cfg, m, om = {"bf16": lambda: bf16_model(LLAMA, 55), "int4": lambda: int4_model(LLAMA, 56),
                  "gqa": lambda: bf16_model(LLAMA_GQA, 61), "mqa": lambda: bf16_model(FALCON_MQA, 62)}[which]()
    prompt = torch.randint(0, cfg.padded_vocab_size, (12,), generator=torch.Generator().manual_seed(7)).to(torch.int32)
    want = O.generate(om, prompt, 112, 112, top_k=1, argmax_ties=True)
    out = lp.generate(m, prompt.to(DEV), 112, 112, top_k=1)
    assert torch.equal(out.cpu(), want), f"first mismatch at {int((out.cpu() != want).nonzero()[0])}"
    assert step_kernel_used(m)
    # sliding window
    om.reset_cache()
    m.reset_cache()
    want = O.generate(om, prompt, 140, 80, top_k=1, argmax_ties=True)
    out = lp.generate(m, prompt.to(DEV), 140, 80, top_k=1)
    assert torch.equal(out.cpu(), want), f"first mismatch at {int((out.cpu() != want).nonzero()[0])}"


def test_step_kernel_matches_per_op_path():
    """Same model, same cache contents: lp_decode_step against the per-op launch sequence (the round-1 path)."""
    cfg, m, _ = bf16_model(NEOX, 57)
    cfg2, m2, _ = bf16_model(NEOX, 57)
    m2.use_step_kernel = False
    g = torch.Generator().manual_seed(9)
    toks = torch.randint(0, cfg.padded_vocab_size, (100,), generator=g)
    pos = torch.arange(70)
    for mm in (m, m2):
        mm._forward_impl(toks[:70].view(1, -1).to(DEV), 128, pos.to(DEV), raw_logits=True)
    for i in range(70, 100):
        p = torch.tensor([i], device=DEV)
        a = m._forward_impl(toks[i].view(1, 1).to(DEV), 128, p, raw_logits=True).float().cpu()
        b = m2._forward_impl(toks[i].view(1, 1).to(DEV), 128, p, raw_logits=True).float().cpu()
        torch.testing.assert_close(a, b, rtol=0, atol=1e-4)
    assert step_kernel_used(m) and not step_kernel_used(m2)
    # the appended rows agree up to the rare bf16 rounding-boundary flip (both paths round fp32 values that differ in the
    # last bits): at most one bf16 ulp, in a handful of elements
    for (ka, va), (kb, vb) in zip(m.kv_caches, m2.kv_caches):
        for a, b in ((ka, kb), (va, vb)):
            a, b = a.float(), b.float()
            assert float((a - b).abs().max()) <= 2.0 ** -7 * float(b.abs().max())
            assert float((a != b).float().mean()) < 1e-3


def test_real_width_llama7b_two_layers_step_kernel():
    """Llama-2-7b widths (E 4096, I 11008, V 32000, hs 128, 32 heads -> 4 sequence splits per head), 2 layers, bf16 weights and
    bf16 cache at a 300-token context: 48 greedy tokens identical to the oracle, logits within the north-star tolerance, and
    within 1e-4 of the per-op path of this library on the same cache contents."""
    cfg = lp.Config.from_name("Llama-2-7b-hf", n_layer=2, block_size=512)
    sd = {k: v.bfloat16() for k, v in O.random_state_dict(cfg, seed=1234).items()}
    models = []
    for step_kernel in (True, False):
        m = lp.GPT(cfg)
        m.load_state_dict(sd)
        m = m.to(device=DEV, dtype=torch.bfloat16).eval()
        m.use_step_kernel = step_kernel
        models.append(m)
    m, m2 = models
    om = O.OracleGPT(cfg, {k: v.float() for k, v in sd.items()}, kv_round=torch.bfloat16)
    prompt = torch.randint(0, cfg.vocab_size, (300,), generator=torch.Generator().manual_seed(1)).to(torch.int32)
    logits = []
    want = O.generate(om, prompt, 348, 348, top_k=1, argmax_ties=True, logits_out=logits)
    out = lp.generate(m, prompt.to(DEV), 348, 348, top_k=1)
    assert torch.equal(out.cpu(), want)
    assert step_kernel_used(m)
    for mm in (m, m2):
        mm.reset_cache()
        mm._forward_impl(prompt.view(1, -1).to(DEV), 348, torch.arange(300, device=DEV), last_only=True, raw_logits=True)
    for i in range(300, 308):
        p = torch.tensor([i], device=DEV)
        a = m._forward_impl(want[i].view(1, 1).to(DEV), 348, p, raw_logits=True)[0, -1].float().cpu()
        b = m2._forward_impl(want[i].view(1, 1).to(DEV), 348, p, raw_logits=True)[0, -1].float().cpu()
        ref = logits[i - 299]
        assert (a - ref).abs().max() < 2e-2 and cosine(a, ref) > 0.999
        # two summation orders of the same fp32 arithmetic (the step kernel applies the RMSNorm scale after the product, the
        # per-op path before it): observed 1.8e-4 on logits of magnitude ~1; both are 100x inside the north-star tolerance above
        torch.testing.assert_close(a, b, rtol=0, atol=4e-4)
    assert not step_kernel_used(m2)


def test_step_kernel_exchange_op_self():
    """The tensor-parallel push exchange of the step kernel on ONE GPU (tp = 1: the only 'peer' is this GPU's own buffer): a
    hand-built op table  W0.x -> slot 0 | x += slot 0 | W1.x -> slot 1 | x += slot 1 | logits = Wl.x , run three times (the
    sender publishes epoch + 1 per use, the receiver waits for it on the local pad, the last exchange of a slot advances the
    epoch counter) with the per-op pull kernel (its own buffer / pad words / state) run in between."""
    import ctypes

    from lit_parrot_b200 import _lib

    lib = _lib.init(0)
    E, V = 512, 64
    g = torch.Generator().manual_seed(77)
    W = [(torch.randn(E, E, generator=g) * 0.05).bfloat16().to(DEV) for _ in range(2)]
    Wl = (torch.randn(V, E, generator=g) * 0.05).bfloat16().to(DEV)
    wte = torch.randn(4, E, generator=g).to(DEV)
    idx = torch.tensor([2], dtype=torch.int32, device=DEV)
    pos = torch.zeros(1, dtype=torch.int32, device=DEV)
    x = torch.zeros(E, device=DEV)
    logits = torch.zeros(V, device=DEV)
    buf = torch.zeros(5 * E, device=DEV)                      # push slots 0 / 1 ([1 rank][E] {value, epoch} pairs each), then a pull slot
    pad = torch.zeros(64, dtype=torch.int32, device=DEV)      # push: words 0, 1 (slot s, rank 0); pull: word 8
    state = torch.zeros(2, 2, dtype=torch.int32, device=DEV)  # push protocol
    pull_state = torch.zeros(2, dtype=torch.int32, device=DEV)
    bufs = torch.tensor([buf.data_ptr()], dtype=torch.int64, device=DEV)
    pads = torch.tensor([pad.data_ptr()], dtype=torch.int64, device=DEV)
    recs = [_lib.LpWeight(w.data_ptr(), None, None, None, None, _lib.LP_W_BF16, w.shape[0], E, 0, 0, 0) for w in (*W, Wl)]
    ops = (_lib.LpStepOp * 5)()

    def tp_fields(o, slot):
        o.tp_buf_ptrs, o.tp_pad_ptrs, o.tp_state = bufs.data_ptr(), pads.data_ptr(), state[slot].data_ptr()
        o.tp_buf_offset, o.tp_pad_base, o.tp_rank, o.tp_size = slot * E * 8, slot, 0, 1

    def lin(i, rec, out, dep, slot=None):
        ops[i].kind, ops[i].dep, ops[i].W, ops[i].x, ops[i].norm_kind = _lib.LP_STEP_LINEAR, dep, ctypes.pointer(rec), x.data_ptr(), -1
        ops[i].epilogue, ops[i].out = _lib.LP_EPI_NONE, out
        if slot is not None:
            tp_fields(ops[i], slot)  # sender of the push exchange that follows

    def exch(i, slot, dep):
        ops[i].kind, ops[i].dep, ops[i].norm_kind = _lib.LP_STEP_EXCHANGE, dep, -1
        tp_fields(ops[i], slot)
        ops[i].residual, ops[i].out = x.data_ptr(), x.data_ptr()

    lin(0, recs[0], buf.data_ptr(), -1, slot=0)
    exch(1, 0, 0)
    lin(2, recs[1], buf.data_ptr() + E * 8, 1, slot=1)
    exch(3, 1, 2)
    lin(4, recs[2], logits.data_ptr(), 3)
    ws = torch.zeros(lib.lp_decode_step_workspace_bytes(1, 64), dtype=torch.uint8, device=DEV)
    gm = _lib.LpStepGeom()
    gm.pos, gm.idx, gm.idx_offset, gm.wte, gm.x0 = pos.data_ptr(), idx.data_ptr(), None, wte.data_ptr(), x.data_ptr()
    gm.cos = gm.sin = None
    gm.workspace, gm.workspace_bytes = ws.data_ptr(), ws.numel()
    gm.idx_is_int64, gm.wte_dtype, gm.E, gm.H, gm.G, gm.hs, gm.n_elem, gm.max_seq, gm.kv_dtype, gm.scale = 0, _lib.LP_F32, E, 1, 1, 64, 0, 64, _lib.LP_BF16, 1.0
    plan = torch.zeros(lib.lp_decode_step_plan_bytes(5), dtype=torch.uint8, device=DEV)
    handle = _lib.LpStepHandle()
    _lib.check(lib.lp_decode_step_plan(ops, 5, ctypes.byref(gm), plan.data_ptr(), plan.numel(), ctypes.byref(handle)), "plan")
    x0 = wte[2].double()
    x1 = x0 + W[0].double() @ x0
    x2 = x1 + W[1].double() @ x1
    want = (Wl.double() @ x2).float()
    stream = torch.cuda.current_stream().cuda_stream
    for rep in range(3):
        _lib.check(lib.lp_decode_step(ctypes.byref(handle), stream), "lp_decode_step")
        torch.cuda.synchronize()
        torch.testing.assert_close(logits, want, rtol=1e-4, atol=1e-4)
        torch.testing.assert_close(x, x2.float(), rtol=1e-4, atol=1e-4)
        # the per-op pull exchange in between (prefill path): disjoint slot, pad word and state
        buf[4 * E:].copy_(torch.arange(E, device=DEV).float())
        tmp = torch.zeros(E, device=DEV)
        _lib.check(lib.lp_tp_allreduce_residual(bufs.data_ptr(), pads.data_ptr(), 0, 1, 4 * E * 4, 8, pull_state.data_ptr(), E, None,
                                                tmp.data_ptr(), 0, stream), "lp_tp_allreduce_residual")
        torch.cuda.synchronize()
        torch.testing.assert_close(tmp, buf[4 * E:])
    assert state[0, 0].item() == 3 and state[1, 0].item() == 3 and pull_state[0].item() == 3
    assert lib.lp_decode_step_status(ctypes.byref(handle), None) == 0


def test_step_kernel_watchdog_reports_instead_of_hanging(monkeypatch):
    """A bound of 50 ns on the cross-CTA waits expires in every real step: the kernel must still run to its end (no hang), the
    sticky error record must say which dependency was seen late, and the host must get LP_ERR_TIMEOUT -> RuntimeError; with the
    default bound the same model then decodes correctly again."""
    import ctypes

    monkeypatch.setenv("LP_DS_TIMEOUT_NS", "50")
    cfg, m, om = bf16_model(LLAMA, 63)
    toks = torch.randint(0, cfg.padded_vocab_size, (24,), generator=torch.Generator().manual_seed(3))
    m._forward_impl(toks[:16].view(1, -1).to(DEV), 64, torch.arange(16, device=DEV), raw_logits=True)
    for i in range(16, 20):
        m._forward_impl(toks[i].view(1, 1).to(DEV), 64, torch.tensor([i], device=DEV), raw_logits=True)
    torch.cuda.synchronize()
    assert step_kernel_used(m)
    eng = m._engine
    with pytest.raises(RuntimeError, match="timed out"):
        eng.check_step_health()
    eng.check_step_health()  # reported once, record cleared
    ent = next(v for v in eng._steps.values() if v is not None)
    assert eng.lib.lp_decode_step_cooperative(ctypes.byref(ent[0])) in (0, 1)
    monkeypatch.delenv("LP_DS_TIMEOUT_NS")
    cfg, m, om = bf16_model(LLAMA, 63)
    teacher_forced(m, om, cfg, prompt_len=20, steps=12, max_seq=64)
    m._engine.check_step_health()


def test_out_of_range_inputs_raise_like_the_reference():
    """Token ids outside the embedding table and positions outside the RoPE table: IndexError on the host-validated paths; on the
    replayed decode graph the device clamps (no out-of-bounds access) and the next call / the health check raises."""
    cfg, m, _ = bf16_model(LLAMA, 64)
    V, bs = cfg.padded_vocab_size, cfg.block_size
    good = torch.randint(0, V, (1, 8)).to(DEV)
    with pytest.raises(IndexError):
        m(torch.full((1, 8), V, device=DEV), 64, torch.arange(8, device=DEV))
    with pytest.raises(IndexError):
        m(good, 64, torch.arange(bs - 4, bs + 4, device=DEV))
    m(good, 64, torch.arange(8, device=DEV))
    m(good[:, :1], 64, torch.tensor([8], device=DEV))            # replayed path, valid
    m(torch.full((1, 1), -3, device=DEV), 64, torch.tensor([9], device=DEV))   # clamped on the device, flagged
    torch.cuda.synchronize()
    with pytest.raises(IndexError, match="token id"):
        m(good[:, :1], 64, torch.tensor([10], device=DEV))
    m(good[:, :1], 64, torch.tensor([10], device=DEV))
    m._engine.check_step_health()


# Llama-style, hs 128, wide enough for the fused column->row pairs (SLAB ops): E / P = 512 rows (one stage pair) per head CTA;
# I = 2816 gives the 148 CTAs 2-3 SwiGLU units each, some of them straddling a 128-column scale group; I = 1024 leaves CTAs
# without any unit; I = 18944 is the 16-unit maximum
LLAMA_SLAB = dict(block_size=256, vocab_size=320, padding_multiple=64, n_layer=3, n_head=16, n_embd=2048, rotary_percentage=1.0,
                  parallel_residual=False, bias=False, _norm_class="RMSNorm", _mlp_class="LLaMAMLP", intermediate_size=2816)


def slab_ops_used(m) -> bool:
    return bool(m._engine._slabs)


@pytest.mark.parametrize("inter", [2816, 1024, 18944], ids=["2-3_units", "some_ctas_empty", "16_units"])
def test_step_kernel_fused_slabs_int4(inter):
    """GPTQ int4 g128 with bf16-exact scales (the storage of a bf16 checkpoint): attention -> attn.proj and fc -> mlp.proj run as
    column->row pairs inside each CTA (SLAB ops).  Teacher-forced logits against the oracle, across key-tile and split boundaries."""
    cfg, m, om = int4_model(dict(LLAMA_SLAB, intermediate_size=inter, n_layer=2 if inter > 4096 else 3), 65, scale_dtype=torch.bfloat16)
    # (the 40-row prompt goes through the 2-term tensor-core GEMM: its error grows with the width, 6e-4 observed at E = 2048)
    teacher_forced(m, om, cfg, prompt_len=40, steps=70, max_seq=256, prompt_atol=2e-3)
    assert step_kernel_used(m) and slab_ops_used(m)
    m._engine.check_step_health()


def test_step_kernel_fused_slabs_match_unfused(monkeypatch):
    """Same weights through the fused (3 grid dependencies per layer) and the five-op table (LP_DS_FUSE=0): greedy tokens identical
    to the oracle on both, logits of the two within 2e-4 of each other."""
    cfg, m, om = int4_model(LLAMA_SLAB, 66, scale_dtype=torch.bfloat16)
    prompt = torch.randint(0, cfg.padded_vocab_size, (12,), generator=torch.Generator().manual_seed(7)).to(torch.int32)
    want = O.generate(om, prompt, 100, 100, top_k=1, argmax_ties=True)
    out = lp.generate(m, prompt.to(DEV), 100, 100, top_k=1)
    assert torch.equal(out.cpu(), want), f"first mismatch at {int((out.cpu() != want).nonzero()[0])}"
    assert slab_ops_used(m)
    monkeypatch.setenv("LP_DS_FUSE", "0")
    cfg2, m2, _ = int4_model(LLAMA_SLAB, 66, scale_dtype=torch.bfloat16)
    out2 = lp.generate(m2, prompt.to(DEV), 100, 100, top_k=1)
    assert torch.equal(out2.cpu(), want) and step_kernel_used(m2) and not slab_ops_used(m2)
    for mm in (m, m2):
        mm.reset_cache()
        mm._forward_impl(want[:60].view(1, -1).to(DEV), 100, torch.arange(60, device=DEV), last_only=True, raw_logits=True)
    for i in range(60, 70):
        p = torch.tensor([i], device=DEV)
        a = m._forward_impl(want[i].view(1, 1).to(DEV), 100, p, raw_logits=True)[0, -1].float().cpu()
        b = m2._forward_impl(want[i].view(1, 1).to(DEV), 100, p, raw_logits=True)[0, -1].float().cpu()
        torch.testing.assert_close(a, b, rtol=0, atol=2e-4)


def bnb_model(kw, mode, seed, cache_bf16=True):
    """bitsandbytes-style weights (NF4 / row-wise int8, restated formats: parity unpinned) + the oracle on the dequantised weights."""
    cfg = lp.Config(**kw) if isinstance(kw, dict) else kw
    fsd = {k: v.bfloat16().float() for k, v in O.random_state_dict(cfg, seed=seed, perturb_norm=True).items()}
    with lp.quantization(mode):
        m = lp.GPT(cfg)
    m.load_state_dict(fsd)
    m = m.to(DEV).eval()
    m.kv_cache_dtype = torch.bfloat16
    dsd = {}
    for k, v in fsd.items():
        if v.dim() == 2 and "wte" not in k:
            dsd[k] = O.nf4_dequantize(*O.nf4_quantize(v), v.shape) if mode == "bnb.nf4" else O.int8_dequantize(*O.int8_quantize(v))
        else:
            dsd[k] = v
    return cfg, m, O.OracleGPT(cfg, dsd, kv_round=torch.bfloat16)


# widths that the streaming formats of the step kernel cover: K % 256 == 0 for NF4 (256 codes per 128-byte K-block)
LLAMA_BNB = dict(LLAMA, n_embd=512, n_head=4, intermediate_size=768)
NEOX_BNB = dict(NEOX, n_embd=512, n_head=4)


@pytest.mark.parametrize("mode", ["bnb.nf4", "bnb.int8"])
@pytest.mark.parametrize("kw,seed", [(LLAMA_BNB, 81), (NEOX_BNB, 82)], ids=["llama", "neox"])
def test_step_kernel_logits_bnb_formats(mode, kw, seed):
    """NF4 (shared-memory table -> bf16 terms -> HMMA) and row-wise int8 (IMMA s8 x s8) through the persistent step kernel."""
    cfg, m, om = bnb_model(kw, mode, seed)
    teacher_forced(m, om, cfg, prompt_len=40, steps=50, max_seq=256, prompt_atol=2e-3)
    assert step_kernel_used(m), "bnb formats did not go through lp_decode_step"
    m._engine.check_step_health()
