"""SURVEY §8 f4 — LLaMA-Adapter, Adapter v2 and merged-LoRA inference against golden data of the UNMODIFIED reference
(lit_gpt/adapter.py, adapter_v2.py, lora.py; fixtures tests/golden/ft_*.npz written by oracle/make_golden.py::finetuned_cases).
CPU: the oracle restatement and the host-side surface (module tree, state-dict keys, filters).  GPU: lp_adapter_attn, the out_bias /
out_scale epilogues and lp_lora_merge through the drop-in models."""
import hashlib
import inspect
import io
import json
import os
from contextlib import redirect_stderr, redirect_stdout
from pathlib import Path

import numpy as np
import pytest
import torch

import lit_parrot_b200 as lp
from lit_parrot_b200 import adapter as lp_adapter
from lit_parrot_b200 import adapter_v2 as lp_adapter_v2
from lit_parrot_b200 import lora as lp_lora
from oracle import lit_oracle as O
from helpers import GOLDEN, cosine, t

FT = ["adapter_neox", "adapter_llama_gqa", "adapter_v2_llama_mha", "adapter_v2_falcon_mqa", "lora_llama_gqa", "lora_neox"]
DEV = "cuda:0"
TIGHT = dict(rtol=0, atol=3e-5)


def digest(sd) -> str:
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].contiguous().numpy().tobytes())
    return h.hexdigest()


def load_case(name):
    z = np.load(os.path.join(GOLDEN, f"ft_{name}.npz"), allow_pickle=False)
    kw = {k: eval(v) for k, v in zip(z["cfg_keys"].tolist(), z["cfg_vals"].tolist())}
    extra = {k: eval(v) for k, v in zip(z["extra_keys"].tolist(), z["extra_vals"].tolist())}
    kind = str(z["kind"])
    cfg = lp.Config(**kw)
    sd = O.random_state_dict(cfg, seed=int(z["seed"]), perturb_norm=True)
    if kind == "lora":
        en = (extra["to_query"], extra["to_key"], extra["to_value"])
        sd.update(O.lora_extra_state(cfg, int(z["seed"]) + 1, extra["r"], en, extra["to_projection"], extra["to_mlp"], extra["to_head"]))
    else:
        sd.update(O.adapter_extra_state(cfg, int(z["seed"]) + 1, extra["adapter_start_layer"], extra["adapter_prompt_length"],
                                        kind == "adapter_v2"))
    assert digest(sd) == str(z["sd_sha256"]), "seeded weights differ from the ones the reference was run on"
    return z, kind, kw, extra, cfg, sd


def oracle_state(kind, cfg, extra, sd):
    if kind != "lora":
        return sd
    return O.lora_merge_state_dict(cfg, sd, extra["r"], extra["alpha"], (extra["to_query"], extra["to_key"], extra["to_value"]))


@pytest.mark.parametrize("name", FT)
def test_oracle_matches_reference_finetuned_golden(name):
    z, kind, kw, extra, cfg, sd = load_case(name)
    om = O.OracleGPT(cfg, oracle_state(kind, cfg, extra, sd))
    idx = t(z["idx"])
    torch.testing.assert_close(om(idx), t(z["ref_full"]), rtol=0, atol=2e-6)
    om.reset_cache()
    max_seq = int(z["max_seq"])
    pos = torch.arange(idx.shape[1])
    torch.testing.assert_close(om(idx, max_seq, pos), t(z["ref_prefill"]), rtol=0, atol=2e-6)
    for s, tok in enumerate(t(z["forced"])):
        pos = pos[-1:] + 1
        torch.testing.assert_close(om(tok, max_seq, pos), t(z["ref_steps"][s]), rtol=0, atol=2e-6)
    om.reset_cache()
    assert torch.equal(O.generate(om, t(z["prompt"]), 30, 30, top_k=1, argmax_ties=True), t(z["ref_gen"]))


def build(kind, kw, extra, sd, dtype=torch.float32, device=None):
    if kind == "lora":
        m = lp_lora.GPT(lp_lora.Config(**kw, **extra))
    else:
        m = lp_adapter.GPT(lp_adapter.Config(**kw, **extra))
        if kind == "adapter_v2":
            lp_adapter_v2.add_adapter_v2_parameters_to_linear_layers(m)
    res = m.load_state_dict(sd, strict=True)  # the reference's keys, nothing missing, nothing unexpected
    assert not res.missing_keys and not res.unexpected_keys
    if device is not None:
        m = m.to(device=device, dtype=dtype)
    return m.eval()


@pytest.mark.parametrize("name", FT)
def test_host_surface_has_the_reference_state_dict(name):
    z, kind, kw, extra, cfg, sd = load_case(name)
    m = build(kind, kw, extra, sd)
    assert set(m.state_dict()) == set(sd)  # e.g. `lora_ind` is not persistent, `adapter_wte` only from adapter_start_layer on
    for k, v in m.state_dict().items():
        assert v.shape == sd[k].shape, k
    if kind == "lora":
        assert all(lp_lora.lora_filter(k, None) == (".lora_" in k) for k in sd)
        lp_lora.mark_only_lora_as_trainable(m)
        assert {n for n, p in m.named_parameters() if p.requires_grad} == {k for k in sd if ".lora_" in k}
        with pytest.raises(RuntimeError, match="GPU"):
            lp_lora.merge_lora_weights(m)  # no CPU path
    elif kind == "adapter":
        keys = {k for k in sd if lp_adapter.adapter_filter(k, None)}
        assert keys == {k for k in sd if "adapter_wte" in k or "gating_factor" in k} and keys
        lp_adapter.mark_only_adapter_as_trainable(m)
        assert {n for n, p in m.named_parameters() if p.requires_grad} == keys
    else:
        keys = {k for k in sd if lp_adapter_v2.adapter_filter(k, None)}
        assert {k for k in sd if "adapter_scale" in k or "norm_" in k or "ln_f" in k} <= keys
        lp_adapter_v2.mark_only_adapter_v2_as_trainable(m)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 4, dtype=torch.long))  # CPU tensors: no CPU path


def test_drop_in_import_paths_and_cli_surface():
    import generate.adapter as ga
    import generate.adapter_v2 as ga2
    import generate.lora as gl
    import lit_gpt.adapter as la
    import lit_gpt.adapter_v2 as la2
    import lit_gpt.lora as ll
    from lit_parrot_b200 import cli_finetuned as cf

    assert la.GPT is lp_adapter.GPT and la2.add_adapter_v2_parameters_to_linear_layers is lp_adapter_v2.add_adapter_v2_parameters_to_linear_layers
    assert ll.merge_lora_weights is lp_lora.merge_lora_weights and ll.LoRAQKVLinear is lp_lora.LoRAQKVLinear
    want = ["prompt", "input", "adapter_path", "checkpoint_dir", "quantize", "max_new_tokens", "top_k", "temperature", "strategy",
            "devices", "precision"]  # generate/adapter.py:23-35, generate/adapter_v2.py:25-37
    assert list(inspect.signature(ga.main).parameters) == want and list(inspect.signature(ga2.main).parameters) == want
    assert list(inspect.signature(gl.main).parameters) == [w if w != "adapter_path" else "lora_path" for w in want]  # generate/lora.py:28-40
    import generate.full as gf

    assert list(inspect.signature(gf.main).parameters) == [w if w != "adapter_path" else "finetuned_path" for w in want]  # generate/full.py:23-35
    with pytest.raises((NotImplementedError, RuntimeError)):  # quantised fully fine-tuned checkpoints: NotImplementedError upstream too
        gf.main("x", "", Path("nope.pth"), Path(GOLDEN) / "ckpt_tiny_llama", "bnb.nf4")
    # scripts/prepare_alpaca.py:141-155
    assert cf.generate_prompt({"instruction": "Do it", "input": ""}) == (
        "Below is an instruction that describes a task. Write a response that appropriately completes the request.\n\n"
        "### Instruction:\nDo it\n\n### Response:")
    assert cf.generate_prompt({"instruction": "Do it", "input": "x"}) == (
        "Below is an instruction that describes a task, paired with an input that provides further context. Write a response that "
        "appropriately completes the request.\n\n### Instruction:\nDo it\n\n### Input:\nx\n\n### Response:")


# ---------------------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("graph", [True, False])
@pytest.mark.parametrize("name", FT)
def test_finetuned_models_match_reference_golden(name, graph):
    z, kind, kw, extra, cfg, sd = load_case(name)
    m = build(kind, kw, extra, sd, device=DEV)
    m.use_cuda_graph = graph
    idx = t(z["idx"]).to(DEV)
    if kind == "lora":
        with pytest.raises(RuntimeError, match="not merged"):
            m(idx)
        lp_lora.merge_lora_weights(m)
        want = oracle_state(kind, cfg, extra, sd)
        for k, v in m.state_dict().items():  # lp_lora_merge == the reference's merge (conv1d / matmul over r <= 12 terms)
            if ".lora_" not in k:
                torch.testing.assert_close(v.cpu(), want[k], rtol=0, atol=1e-7, msg=k)
        lp_lora.merge_lora_weights(m)  # idempotent (`merged` flag, lora.py:160-164)
        torch.testing.assert_close(m.state_dict()["transformer.h.0.attn.attn.weight"].cpu(), want["transformer.h.0.attn.attn.weight"],
                                   rtol=0, atol=1e-7)
    torch.testing.assert_close(m(idx).cpu(), t(z["ref_full"]), **TIGHT)
    max_seq = int(z["max_seq"])
    pos = torch.arange(idx.shape[1], device=DEV)
    torch.testing.assert_close(m(idx, max_seq, pos).cpu(), t(z["ref_prefill"]), **TIGHT)
    for s, tok in enumerate(t(z["forced"])):
        pos = pos[-1:] + 1
        torch.testing.assert_close(m(tok.to(DEV), max_seq, pos).cpu(), t(z["ref_steps"][s]), **TIGHT)
    if kind != "lora":
        # the reference's adapter_kv_caches (adapter.py:103-107), compact: one head per query group
        caches = m.adapter_kv_caches
        start = extra["adapter_start_layer"]
        assert len(caches) == cfg.n_layer and all(c is None for c in caches[:start]) and all(c is not None for c in caches[start:])
        assert caches[start][0].shape == (1, cfg.n_query_groups, extra["adapter_prompt_length"], cfg.head_size)
    m.reset_cache()
    prompt = t(z["prompt"]).to(DEV)
    out = lp.generate(m, prompt, 30, 30, temperature=1.0, top_k=1)
    assert torch.equal(out.cpu(), t(z["ref_gen"]))
    if kind == "lora":  # a merged model is a plain GPT: batch-1 decode runs through the persistent step kernel where it applies
        assert m._engine.use_step_kernel
    else:
        assert not any(v is not None for v in m._engine._steps.values())


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["adapter_v2_llama_mha", "adapter_llama_gqa", "lora_llama_gqa"])
def test_finetuned_bf16_weights_and_bf16_faithful_mode(name):
    """bf16 parameters: (a) fp32 activations over the bf16-stored weights == the oracle in fp32 on the same rounded weights;
    (b) precision='bf16' == the oracle run in bf16 (the reference's bf16-true arithmetic) within bf16 noise."""
    z, kind, kw, extra, cfg, sd = load_case(name)
    sd16 = {k: v.bfloat16() for k, v in sd.items()}
    m = build(kind, kw, extra, sd16, dtype=torch.bfloat16, device=DEV)
    if kind == "lora":
        lp_lora.merge_lora_weights(m)  # bf16 merge: (B @ A) -> bf16, * scaling -> bf16, += -> bf16
    osd = oracle_state(kind, cfg, extra, sd16)
    if kind == "lora":
        for k, v in m.state_dict().items():
            if ".lora_" not in k:
                d = (v.float().cpu() - osd[k].float()).abs().max().item()
                assert d <= 2.0 ** -8 * osd[k].float().abs().max().item(), (k, d)  # at most one bf16 ulp (dot-product order)
        osd = {k: v.cpu() for k, v in m.state_dict().items() if ".lora_" not in k}  # compare the forward on identical weights
    idx = t(z["idx"]).to(DEV)
    want = O.OracleGPT(cfg, {k: v.float() for k, v in osd.items()})(idx.cpu())
    torch.testing.assert_close(m(idx).float().cpu(), want, rtol=0, atol=2e-2)
    pos = torch.arange(idx.shape[1], device=DEV)
    got = m._forward_impl(idx, 32, pos, raw_logits=True).cpu()  # the engine's fp32 logits (cached prefill == full forward)
    # the cached forward keeps k / v in a bf16 cache like the reference's bf16 run: same rounding in the oracle (kv_round)
    want_c = O.OracleGPT(cfg, {k: v.float() for k, v in osd.items()}, kv_round=torch.bfloat16)(idx.cpu(), 32, pos.cpu())
    torch.testing.assert_close(got, want_c, rtol=0, atol=1e-4)
    m.reset_cache()
    m.set_precision("bf16")
    got16 = m(idx).float().cpu()
    want16 = O.OracleGPT(cfg, osd, dtype=torch.bfloat16)(idx.cpu()).float()
    assert (got16 - want16).abs().max().item() < 6e-2 and (got16 - want).abs().max().item() < 8e-2


@pytest.mark.gpu
def test_adapter_attn_kernel_against_sdpa():
    """lp_adapter_attn alone: partial rotary, GQA, hs 128, T > 1, against the fp64 restatement of adapter.py:251-254."""
    import math

    from lit_parrot_b200 import _lib

    lib = _lib.init(0)
    g = torch.Generator().manual_seed(5)
    B, T, H, G, hs, n_elem, aT, block = 2, 5, 8, 2, 128, 32, 10, 64
    qpk = H // G
    qkv = torch.randn(B * T, (H + 2 * G) * hs, generator=g)
    ak, av = torch.randn(G, aT, hs, generator=g), torch.randn(G, aT, hs, generator=g)
    gate = torch.randn(H, generator=g)
    y0 = torch.randn(B * T, H * hs, generator=g)
    cos, sin = O.rope_tables(block, n_elem, torch.float32)
    pos = torch.tensor([3, 4, 9, 20, 63], dtype=torch.int32)
    q = qkv.view(B, T, G, qpk + 2, hs)[:, :, :, :qpk].reshape(B, T, H, hs).permute(0, 2, 1, 3).double()
    c, s = cos[pos.long()].double(), sin[pos.long()].double()
    q = torch.cat((O.rotate(q[..., :n_elem], c, s), q[..., n_elem:]), dim=-1)
    k = ak.double().repeat_interleave(qpk, dim=0)[None]
    v = av.double().repeat_interleave(qpk, dim=0)[None]
    att = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(hs), dim=-1) @ v  # (B, H, T, hs)
    want = y0.double() + (gate.double().view(1, H, 1, 1) * att).permute(0, 2, 1, 3).reshape(B * T, H * hs)
    d = lambda x: x.to(DEV).contiguous()  # noqa: E731
    out = d(y0)
    bufs = [d(x) for x in (qkv, cos, sin, pos, ak, av, gate)]
    _lib.check(lib.lp_adapter_attn(*(b.data_ptr() for b in bufs), out.data_ptr(), B, T, H, G, hs, n_elem, aT, 1.0 / math.sqrt(hs), 0,
                                   torch.cuda.current_stream().cuda_stream))
    torch.testing.assert_close(out.cpu().double(), want, rtol=0, atol=2e-5)
    assert lib.lp_adapter_attn(*(b.data_ptr() for b in bufs), out.data_ptr(), B, T, H, G, 512, n_elem, aT, 1.0, 0, None) == -2


@pytest.mark.gpu
def test_generate_cli_mains_on_tiny_checkpoint(tmp_path):
    """`python generate/adapter_v2.py` / `generate/lora.py` on the tiny checkpoint + a fine-tuned file written here: the printed
    response is what the oracle decodes from the same files (Alpaca prompt, cut at '### Response:').  The word-level test
    vocabulary gets the words of the Alpaca prompt so that '### Response:' survives the encode / decode round trip."""
    import shutil

    from lit_parrot_b200 import cli_finetuned as cf
    from lit_parrot_b200.tokenizer import Tokenizer

    ckpt = tmp_path / "ckpt"
    shutil.copytree(Path(GOLDEN) / "ckpt_tiny_llama", ckpt)
    words = sorted(set(cf.generate_prompt({"instruction": "", "input": ""}).split()))
    tj = json.load(open(ckpt / "tokenizer.json"))
    by_id = {v: k for k, v in tj["model"]["vocab"].items()}
    for i, w in enumerate(words):
        by_id[60 + i] = w  # replaces filler words w60..
    tj["model"]["vocab"] = {w: i for i, w in by_id.items()}
    json.dump(tj, open(ckpt / "tokenizer.json", "w"))
    cfg = lp.Config(**json.load(open(ckpt / "lit_config.json")))
    base_sd = torch.load(ckpt / "lit_model.pth")
    tok = Tokenizer(ckpt)
    prompt_ids = tok.encode(cf.generate_prompt({"instruction": "w20 w21", "input": ""}))
    assert "### Response:" in tok.decode(prompt_ids)
    for kind in ("adapter_v2", "lora", "full"):  # (Config defaults start the v1 prefix at layer 2: a 2-layer checkpoint has none)
        if kind == "full":  # a fully fine-tuned checkpoint: every weight in the one file (here: the base weights perturbed)
            gfull = torch.Generator().manual_seed(3)
            extra = {k: (v + 0.02 * torch.randn(v.shape, generator=gfull)) if v.dim() == 2 else v.clone() for k, v in base_sd.items()}
        elif kind == "lora":
            extra = O.lora_extra_state(cfg, 7, cf.lora_r, (True, False, True), False, False, False)
        else:
            extra = O.adapter_extra_state(cfg, 7, 2, 10, True)
        path = tmp_path / f"{kind}.pth"
        torch.save({"model": extra} if kind == "lora" else extra, path)
        out, err = io.StringIO(), io.StringIO()
        main = {"adapter_v2": cf.main_adapter_v2, "lora": cf.main_lora, "full": cf.main_full}[kind]
        with redirect_stdout(out), redirect_stderr(err):
            main("w20 w21", "", path, ckpt, None, 12, 1, 1.0, "auto", 1, "32-true")
        sd = dict(base_sd)
        sd.update(extra)
        if kind == "lora":
            sd = O.lora_merge_state_dict(cfg, sd, cf.lora_r, cf.lora_alpha, (True, False, True))
        sd = {k: v.float() for k, v in sd.items()}
        n = prompt_ids.numel() + 12
        want = O.generate(O.OracleGPT(cfg, sd), prompt_ids, n, n, top_k=1, eos_id=tok.eos_id, argmax_ties=True)
        assert out.getvalue().strip() == tok.decode(want).split("### Response:")[1].strip(), kind
        plain = O.generate(O.OracleGPT(cfg, base_sd), prompt_ids, n, n, top_k=1, eos_id=tok.eos_id, argmax_ties=True)
        assert not torch.equal(plain, want), "the fine-tuned weights must change the continuation for this test to mean anything"
        assert "Time for inference:" in err.getvalue() and "Memory used:" in err.getvalue()


@pytest.mark.gpu
@pytest.mark.parametrize("kind,preset", [("adapter_v2", "Llama-2-7b-hf"), ("adapter", "stablelm-base-alpha-3b")])
def test_adapter_models_at_real_widths(kind, preset):
    """The adapter paths at the widths of the benchmarked models (two layers deep, bf16 weights, fp32 activations): Llama-2-7b with
    Adapter v2 (out_bias / out_scale through the tcgen05 GEMM on 24 prompt rows and the streaming GEMV on the decode rows, SwiGLU
    vectors interleaved, 11008-wide MLP) and stablelm-3b with the v1 prefix (32 heads of 128, 25 % rotary, LayerNorm + biases,
    V 50688): cached prefill + teacher-forced decode steps against the oracle on the same bf16-rounded weights and bf16 cache."""
    kw = dict(lp.name_to_config[preset], n_layer=2, block_size=64)
    extra = dict(adapter_prompt_length=10, adapter_start_layer=0)
    cfg = lp.Config(**kw)
    sd = {k: v.bfloat16() for k, v in O.random_state_dict(cfg, seed=21, perturb_norm=True).items()}
    sd.update({k: v.bfloat16() for k, v in O.adapter_extra_state(cfg, 22, 0, 10, kind == "adapter_v2").items()})
    m = build(kind, kw, extra, sd, dtype=torch.bfloat16, device=DEV)
    om = O.OracleGPT(cfg, {k: v.float() for k, v in sd.items()}, kv_round=torch.bfloat16)
    g = torch.Generator().manual_seed(4)
    idx = torch.randint(0, cfg.vocab_size, (1, 24), generator=g)
    pos = torch.arange(24)
    got = m._forward_impl(idx.to(DEV), 64, pos.to(DEV), raw_logits=True).cpu()
    want = om(idx, 64, pos)
    # tolerance of the real-shape tests (test_real_shapes_gpu.py): 2e-3 of the logit scale — k / v are rounded to bf16 on append on
    # both sides, and a last-bit difference of an fp32 value flips that rounding now and then
    scale = want.abs().max().item()
    tol = 2e-3 * max(scale, 1.0)
    torch.testing.assert_close(got, want, rtol=0, atol=tol)
    assert cosine(got, want) > 0.99999
    for s in range(4):
        tok = torch.randint(0, cfg.vocab_size, (1, 1), generator=g)
        pos = pos[-1:] + 1
        got = m._forward_impl(tok.to(DEV), 64, pos.to(DEV), raw_logits=True).cpu()
        torch.testing.assert_close(got, om(tok, 64, pos), rtol=0, atol=tol)
    plain = O.OracleGPT(cfg, {k: v.float() for k, v in sd.items() if "adapter" not in k and "gating" not in k}, kv_round=torch.bfloat16)
    assert (plain(idx, 64, torch.arange(24)) - want).abs().max().item() > 10 * tol  # the adapter matters
