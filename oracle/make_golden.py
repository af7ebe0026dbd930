"""Generate tests/golden/*.npz from the UNMODIFIED reference and pin the oracle restatement to it.

Runs only in the build container (needs /root/reference):

    python oracle/make_golden.py

It imports the reference's `lit_gpt`, `generate.base` and `quantize.gptq` through the stubs in
oracle/_shims (lightning / lightning_utilities / nltk are not installed), drives them on seeded
random-init weights, stores inputs + reference outputs as small fixtures, and asserts that
oracle/lit_oracle.py reproduces every stored output (bit-for-bit where the op order is identical).
The GPU box has no /root/reference: tests there only read the committed fixtures.
"""
import hashlib
import math
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = "/root/reference"
# the reference's `lit_gpt`/`generate`/`quantize` must win over this repo's drop-in packages
sys.path = [p for p in sys.path if os.path.abspath(p or ".") != REPO]
sys.path[:0] = [os.path.join(HERE, "_shims"), REF]

import numpy as np  # noqa: E402
import torch  # noqa: E402

import generate.base as ref_generate  # noqa: E402  (reference)
import lit_gpt  # noqa: E402  (reference)
import quantize.gptq as ref_gptq  # noqa: E402  (reference)

assert lit_gpt.__file__.startswith(REF), lit_gpt.__file__

import importlib.util  # noqa: E402

_spec = importlib.util.spec_from_file_location("lit_oracle", os.path.join(HERE, "lit_oracle.py"))
oracle = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(oracle)

OUT = os.path.join(REPO, "tests", "golden")
os.makedirs(OUT, exist_ok=True)

TINY = {
    # NeoX family: parallel residual, two LayerNorms, biases, partial rotary 25 %, MHA
    "neox": dict(block_size=48, vocab_size=96, padding_multiple=32, n_layer=2, n_head=4, n_embd=64,
                 rotary_percentage=0.25, parallel_residual=True, bias=True),
    # RedPajama-like: NeoX with sequential residual and full rotary
    "neox_seq": dict(block_size=48, vocab_size=96, padding_multiple=32, n_layer=2, n_head=4, n_embd=64,
                     rotary_percentage=1.0, parallel_residual=False, bias=True),
    # Falcon-7b-like: MQA, one shared LayerNorm, parallel residual, no linear biases
    "falcon_mqa": dict(block_size=48, padded_vocab_size=80, n_layer=2, n_head=5, n_embd=80, rotary_percentage=1.0,
                       parallel_residual=True, n_query_groups=1, bias=False, shared_attention_norm=True),
    # Falcon-40b-like: GQA, two norms, parallel residual
    "falcon_gqa": dict(block_size=48, padded_vocab_size=80, n_layer=2, n_head=8, n_embd=128, rotary_percentage=1.0,
                       parallel_residual=True, n_query_groups=2, bias=False),
    # Llama-2-7b-like: RMSNorm, SwiGLU, sequential residual, MHA
    "llama_mha": dict(block_size=48, vocab_size=96, padding_multiple=32, n_layer=2, n_head=4, n_embd=64,
                      rotary_percentage=1.0, parallel_residual=False, bias=False, _norm_class="RMSNorm",
                      norm_eps=1e-5, _mlp_class="LLaMAMLP", intermediate_size=176),
    # Llama-2-70b-like: GQA with 2 groups
    "llama_gqa": dict(block_size=48, vocab_size=96, padding_multiple=32, n_layer=3, n_head=8, n_embd=128,
                      n_query_groups=2, rotary_percentage=1.0, parallel_residual=False, bias=False,
                      _norm_class="RMSNorm", norm_eps=1e-5, _mlp_class="LLaMAMLP", intermediate_size=352),
    # LongChat-like: condense_ratio
    "llama_condense": dict(block_size=64, vocab_size=96, padding_multiple=32, n_layer=1, n_head=2, n_embd=64,
                           rotary_percentage=1.0, parallel_residual=False, bias=False, _norm_class="RMSNorm",
                           norm_eps=1e-6, _mlp_class="LLaMAMLP", intermediate_size=96, condense_ratio=8),
}


def sd_digest(sd) -> str:
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].contiguous().numpy().tobytes())
    return h.hexdigest()


def build_reference(cfg_kwargs, seed, perturb=True):
    cfg = lit_gpt.Config(**cfg_kwargs)
    model = lit_gpt.GPT(cfg)
    sd = oracle.random_state_dict(cfg, seed=seed, perturb_norm=perturb)
    missing = model.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    model.eval()
    return cfg, model, sd


def check(name, a, b, atol):
    d = (a.float() - b.float()).abs().max().item()
    assert d <= atol, f"{name}: oracle differs from reference by {d}"
    return d


@torch.no_grad()
def tiny_cases():
    for name, kw in TINY.items():
        seed = 1234
        cfg, model, sd = build_reference(kw, seed)
        g = torch.Generator().manual_seed(1)
        B, T, steps = 2, 7, 12
        idx = torch.randint(0, cfg.padded_vocab_size, (B, T), generator=g)
        max_seq = 32
        # 1) no-cache forward
        ref_full = model(idx)
        # 2) cached prefill + teacher-forced decode steps (B=2 lock-step, model.py:66,131)
        model.reset_cache()
        pos = torch.arange(T)
        ref_prefill = model(idx, max_seq, pos)
        forced = torch.randint(0, cfg.padded_vocab_size, (steps, B, 1), generator=g)
        ref_steps = []
        for s in range(steps):
            pos = pos[-1:] + 1
            ref_steps.append(model(forced[s], max_seq, pos))
        ref_steps = torch.stack(ref_steps)
        ref_k = torch.stack([kv[0] for kv in model.kv_caches])
        ref_v = torch.stack([kv[1] for kv in model.kv_caches])
        # 3) greedy generate (top_k=1 -> a single non-zero probability unless exact ties) incl. the
        #    sliding-window overflow branch (max_seq_length < tokens generated, model.py:238-242)
        model.reset_cache()
        prompt = torch.randint(0, cfg.padded_vocab_size, (5,), generator=g, dtype=torch.int64).to(torch.int32)
        ref_gen = ref_generate.generate(model, prompt, 30, 30, temperature=1.0, top_k=1)
        model.reset_cache()
        ref_gen_overflow = ref_generate.generate(model, prompt, 30, 12, temperature=1.0, top_k=1)
        model.reset_cache()

        # ---- oracle must reproduce all of it
        om = oracle.OracleGPT(oracle_cfg(kw), sd)
        d1 = check(name + "/full", om(idx), ref_full, 0.0)
        om.reset_cache()
        pos = torch.arange(T)
        d2 = check(name + "/prefill", om(idx, max_seq, pos), ref_prefill, 0.0)
        for s in range(steps):
            pos = pos[-1:] + 1
            check(name + f"/step{s}", om(forced[s], max_seq, pos), ref_steps[s], 0.0)
        check(name + "/kcache", torch.stack([kv[0] for kv in om.kv]), ref_k, 0.0)
        om.reset_cache()
        og = oracle.generate(om, prompt, 30, 30, temperature=1.0, top_k=1, argmax_ties=True)
        assert torch.equal(og, ref_gen), (name, og, ref_gen)
        om.reset_cache()
        og2 = oracle.generate(om, prompt, 30, 12, temperature=1.0, top_k=1, argmax_ties=True)
        assert torch.equal(og2, ref_gen_overflow), (name, og2, ref_gen_overflow)
        print(f"[golden] {name}: oracle == reference (max diff {max(d1, d2):.1e}); tokens equal")

        G = cfg.n_query_groups
        qpk = cfg.n_head // G
        # compact view of the reference cache (one head per query group) for the product's layout
        kc = ref_k if G == 1 else ref_k[:, :, ::qpk]
        vc = ref_v if G == 1 else ref_v[:, :, ::qpk]
        np.savez_compressed(
            os.path.join(OUT, f"tiny_{name}.npz"),
            cfg_keys=np.array(list(kw.keys())), cfg_vals=np.array([repr(v) for v in kw.values()]),
            seed=seed, sd_sha256=sd_digest(sd), idx=idx.numpy(), max_seq=max_seq, forced=forced.numpy(),
            ref_full=ref_full.numpy(), ref_prefill=ref_prefill.numpy(), ref_steps=ref_steps.numpy(),
            ref_k_compact=kc.numpy(), ref_v_compact=vc.numpy(), prompt=prompt.numpy(), ref_gen=ref_gen.numpy(),
            ref_gen_overflow=ref_gen_overflow.numpy(),
        )


def oracle_cfg(kw):
    """The oracle only needs attribute access; reuse the product's Config (same fields) without importing
    the product's CUDA side."""
    spec = importlib.util.spec_from_file_location("lp_config", os.path.join(REPO, "lit_parrot_b200", "config.py"))
    m = importlib.util.module_from_spec(spec)
    sys.modules["lp_config"] = m
    spec.loader.exec_module(m)
    return m.Config(**kw)


@torch.no_grad()
def preset_table():
    """Every preset: derived values from the reference Config, to pin lit_parrot_b200.config."""
    rows = {}
    for name in lit_gpt.config.name_to_config:
        c = lit_gpt.Config.from_name(name)
        rows[name] = [c.block_size, c.vocab_size, c.padded_vocab_size, c.n_layer, c.n_head, c.n_embd,
                      c.n_query_groups, c.intermediate_size, int(c.rotary_percentage * 1000), int(c.parallel_residual),
                      int(c.bias), int(c.shared_attention_norm), int(c._norm_class == "RMSNorm"),
                      int(c._mlp_class == "LLaMAMLP"), c.condense_ratio, int(round(c.norm_eps * 1e9)), c.head_size]
    names = sorted(rows)
    np.savez_compressed(os.path.join(OUT, "presets.npz"), names=np.array(names),
                        table=np.array([rows[n] for n in names], dtype=np.int64),
                        orgs=np.array([lit_gpt.config.name_to_config[n]["org"] for n in names]))
    print(f"[golden] presets: {len(names)} configs")


@torch.no_grad()
def pythia70m():
    """BASELINE config 1: pythia-70m random init, fp32, greedy generate 16 -> 128 tokens on CPU."""
    cfg = lit_gpt.Config.from_name("pythia-70m")
    model = lit_gpt.GPT(cfg)
    sd = oracle.random_state_dict(cfg, seed=1234)
    model.load_state_dict(sd)
    model.eval()
    prompt = torch.randint(0, cfg.vocab_size, (16,), generator=torch.Generator().manual_seed(1)).to(torch.int32)
    toks = ref_generate.generate(model, prompt, 128, 128, temperature=1.0, top_k=1)
    # logits of the reference at every decode step (teacher-forced by its own tokens), last row only
    model.reset_cache()
    pos = torch.arange(16)
    lg = [model(toks[:16].view(1, -1).long(), 128, pos)[0, -1]]
    for t in range(16, 127):
        pos = pos[-1:] + 1
        lg.append(model(toks[t].view(1, 1).long(), 128, pos)[0, -1])
    lg = torch.stack(lg)  # (112, V)
    top2 = torch.topk(lg, 2, dim=-1).values
    gaps = (top2[:, 0] - top2[:, 1])
    assert torch.equal(lg.argmax(-1).to(torch.int32), toks[16:])
    om = oracle.OracleGPT(oracle_cfg(dict(lit_gpt.config.name_to_config["pythia-70m"])), sd)
    otoks = oracle.generate(om, prompt, 128, 128, temperature=1.0, top_k=1, argmax_ties=True)
    assert torch.equal(otoks, toks)
    # keep the fixture small: logits at 8 probe steps + per-step top-8 (ids + values) + gaps
    probe = [0, 1, 2, 15, 31, 63, 95, 111]
    t8 = torch.topk(lg, 8, dim=-1)
    np.savez_compressed(os.path.join(OUT, "pythia70m_greedy.npz"), seed=1234, prompt=prompt.numpy(),
                        tokens=toks.numpy(), probe_steps=np.array(probe), probe_logits=lg[probe].numpy(),
                        top8_ids=t8.indices.numpy().astype(np.int32), top8_vals=t8.values.numpy(),
                        gaps=gaps.numpy(), sd_sha256=sd_digest(sd))
    print(f"[golden] pythia-70m: 112 greedy tokens, min top1-top2 gap {gaps.min():.3e}; oracle tokens equal")


@torch.no_grad()
def gptq_cases():
    """ColBlockQuantizedLinear storage + dequant + forward (quantize/gptq.py:205-264), g=128 and per-row."""
    g = torch.Generator().manual_seed(7)
    for tag, tile in (("g128", 128), ("perrow", -1)):
        out_f, in_f = 48, 384
        w = torch.randn(out_f, in_f, generator=g) * 0.02
        lin = ref_gptq.ColBlockQuantizedLinear(in_f, out_f, True, bits=4, tile_cols=tile)
        packed, scales, zeros = oracle.gptq_rtn_quantize(w, tile)
        # cross-check the grid against the reference's own find_params_weight on each tile
        dummy = torch.nn.Linear(in_f, out_f)
        qz = ref_gptq.GPTQQuantizer(dummy, bits=4, groupsize=tile)
        tc = in_f if tile == -1 else tile
        for j in range(scales.shape[1]):
            s, z = qz.find_params_weight(w[:, j * tc:(j + 1) * tc])
            assert torch.equal(s, scales[:, j:j + 1]) and torch.equal(z, zeros[:, j:j + 1])
        lin.scales.copy_(scales)
        lin.zeros.copy_(zeros)
        lin.bias.copy_(torch.randn(out_f, generator=g) * 0.02)
        # the reference packer applied to the on-grid weights must give the same bytes
        wq = oracle.gptq_dequant(packed, scales, zeros)
        lin.pack_weight(wq + 1e-4 * scales.repeat_interleave(tc, dim=1)[:, :in_f])  # nudge off truncation edges
        assert torch.equal(lin.quant_weight, packed), "nibble packing differs from reference pack_weight"
        assert lin.quant_weight.stride() == (1, out_f)
        ref_w = lin.get_weight(torch.float32)
        check(f"gptq/{tag}/dequant", wq, ref_w, 0.0)
        x = torch.randn(3, in_f, generator=g)
        y = lin(x)
        oy = oracle.linear(x, {"l.quant_weight": packed, "l.scales": scales, "l.zeros": zeros, "l.bias": lin.bias}, "l")
        check(f"gptq/{tag}/forward", oy, y, 0.0)
        ref_w_bf16 = lin.get_weight(torch.bfloat16)
        check(f"gptq/{tag}/dequant_bf16", oracle.gptq_dequant(packed, scales, zeros, dtype=torch.bfloat16), ref_w_bf16, 0.0)
        np.savez_compressed(os.path.join(OUT, f"gptq_{tag}.npz"), w=w.numpy(), quant_weight=packed.contiguous().numpy(),
                            scales=scales.numpy(), zeros=zeros.numpy(), bias=lin.bias.numpy(), x=x.numpy(), y=y.numpy(),
                            dequant=ref_w.numpy(), dequant_bf16_as_f32=ref_w_bf16.float().numpy(), tile_cols=tile)
        print(f"[golden] gptq {tag}: packing, dequant (fp32+bf16) and forward equal to reference")

    # the operator plug-in switch builds a fully quantised GPT (lit_gpt/utils.py:26-83)
    from lit_gpt.utils import quantization

    with quantization("gptq.int4"):
        m = lit_gpt.GPT(lit_gpt.Config(**TINY["llama_mha"]))
    keys = sorted(m.state_dict().keys())
    assert torch.nn.Linear is not ref_gptq.ColBlockQuantizedLinear
    np.savez_compressed(os.path.join(OUT, "gptq_statedict_keys.npz"), keys=np.array(keys),
                        shapes=np.array([repr(tuple(m.state_dict()[k].shape)) for k in keys]))
    print(f"[golden] gptq plug-in: {len(keys)} state-dict keys recorded")


@torch.no_grad()
def checkpoint_cases():
    """Checkpoint directories as the reference's own tools leave them: `lit_config.json` =
    json.dump(config.__dict__) (scripts/convert_hf_checkpoint.py:193-194), `lit_model.pth` = the GPT state dict,
    `lit_model_gptq.4bit.pth` = state dict of the model built under quantization("gptq.int4") (quantize/gptq.py:595-596).
    Read back with the reference's lazy_load + load_state_dict (generate/base.py:219-221); the logits of that reloaded
    reference model are the expected outputs stored next to the files."""
    import json
    from pathlib import Path

    from lit_gpt.utils import lazy_load, quantization

    # intermediate_size 192: the int4 kernels need in_features % 32 == 0 (true of every preset; 176 is not)
    kw = dict(TINY["llama_mha"], intermediate_size=192)
    cfg, model, sd = build_reference(kw, seed=4321)
    d = Path(OUT) / "ckpt_tiny_llama"
    d.mkdir(exist_ok=True)
    with open(d / "lit_config.json", "w") as fp:
        json.dump(cfg.__dict__, fp)
    torch.save(model.state_dict(), d / "lit_model.pth")

    # int4 file: every Linear (lm_head included, utils.py:72-74: per-row groups) round-to-nearest quantised
    with quantization("gptq.int4"):
        qmodel = lit_gpt.GPT(cfg)
    qsd = {}
    for k, v in sd.items():
        if k.endswith(".weight") and v.dim() == 2 and "wte" not in k:
            packed, scales, zeros = oracle.gptq_rtn_quantize(v.float(), -1)
            base = k[: -len("weight")]
            qsd[base + "quant_weight"], qsd[base + "scales"], qsd[base + "zeros"] = packed, scales, zeros
        else:
            qsd[k] = v
    res = qmodel.load_state_dict(qsd, strict=False)
    assert not res.unexpected_keys, res.unexpected_keys
    assert qmodel.lm_head.quant_weight.stride() == (1, cfg.padded_vocab_size)
    torch.save(qmodel.state_dict(), d / "lit_model_gptq.4bit.pth")

    g = torch.Generator().manual_seed(11)
    idx = torch.randint(0, cfg.vocab_size, (2, 9), generator=g)
    out = {}
    for tag, fname, quant in (("fp32", "lit_model.pth", None), ("int4", "lit_model_gptq.4bit.pth", "gptq.int4")):
        with quantization(quant):
            m = lit_gpt.GPT(cfg)
        with lazy_load(d / fname) as ckpt:
            m.load_state_dict(ckpt.get("model", ckpt), strict=quant is None)
        m.eval()
        out[tag] = m(idx).float().numpy()
        digest = sd_digest({k: (v.contiguous() if torch.is_tensor(v) else v) for k, v in m.state_dict().items()})
        out[tag + "_digest"] = np.array(digest)
    # the oracle on the same files (plain torch.load)
    o32 = oracle.OracleGPT(oracle_cfg(kw), torch.load(d / "lit_model.pth"))(idx)
    o4 = oracle.OracleGPT(oracle_cfg(kw), torch.load(d / "lit_model_gptq.4bit.pth"))(idx)
    check("ckpt/int4/oracle", o4, torch.from_numpy(out["int4"]), 1e-5)
    check("ckpt/fp32/oracle", o32, torch.from_numpy(out["fp32"]), 1e-5)
    np.savez_compressed(os.path.join(OUT, "ckpt_tiny_llama_expected.npz"), idx=idx.numpy(), logits_fp32=out["fp32"],
                        logits_int4=out["int4"], digest_fp32=out["fp32_digest"], digest_int4=out["int4_digest"])
    print(f"[golden] checkpoint dir {d.name}: fp32 + gptq.int4 files written, reference reload logits stored")


def cli_cases():
    """SURVEY §8 f3 — the callers on the other side of generate(): the reference's Tokenizer (lit_gpt/tokenizer.py), the streaming
    chat generate with stop sequences (chat/base.py:20-95) and prompt_config (chat/base.py:202-290), driven here UNMODIFIED:
    * a tiny HF-tokenizers vocabulary (word level, 96 entries = the tiny checkpoint's vocab) is written into ckpt_tiny_llama/, which
      thereby becomes a complete checkpoint directory (check_valid_checkpoint_dir passes);
    * encode / decode round trips, the token streams chat.generate yields for scripted next-token sequences (a stub model whose
      logits are one-hot, so multinomial is deterministic), and the (template, stop ids) pairs of every model family are stored
      in tests/golden/cli_golden.json."""
    import json
    from pathlib import Path

    from tokenizers import Tokenizer as HFTokenizer
    from tokenizers.models import WordLevel
    from tokenizers.pre_tokenizers import WhitespaceSplit

    import chat.base as ref_chat  # reference
    from lit_gpt import Tokenizer  # reference

    assert ref_chat.__file__.startswith(REF), ref_chat.__file__
    d = Path(OUT) / "ckpt_tiny_llama"
    words = ["<|endoftext|>", "<|SYSTEM|>", "<|ASSISTANT|>", "<|USER|>", "<", ">:", "human", "bot", "Q", ":", "Question", "A",
             "Label", "User", "[UNK]"]
    words += [f"w{i}" for i in range(len(words), 96)]
    tok = HFTokenizer(WordLevel({w: i for i, w in enumerate(words)}, unk_token="[UNK]"))
    tok.pre_tokenizer = WhitespaceSplit()
    tok.save(str(d / "tokenizer.json"))
    with open(d / "tokenizer_config.json", "w") as fp:
        json.dump({"bos_token": "<|endoftext|>", "eos_token": "<|endoftext|>"}, fp)
    t = Tokenizer(d)
    gold = {"vocab_size": t.vocab_size, "bos_id": t.bos_id, "eos_id": t.eos_id, "backend": t.backend, "encode": [], "decode": []}
    for text, kw in (("w20 w21 human : w95", {}), ("Q : w30 unknownword A :", {"bos": True}), ("w40 w41 w42 w43", {"eos": True}),
                     ("w50 w51 w52 w53 w54", {"bos": True, "eos": True, "max_length": 4}), ("", {})):
        ids = t.encode(text, **kw)
        gold["encode"].append({"text": text, "kw": kw, "ids": ids.tolist(), "dtype": str(ids.dtype)})
    for ids in ([20, 21, 6, 9, 95], [0], [33]):
        gold["decode"].append({"ids": ids, "text": t.decode(torch.tensor(ids))})
    gold["decode"].append({"ids": 33, "text": t.decode(torch.tensor(33))})  # 0-dim tensor

    class Scripted(torch.nn.Module):
        """next-token script: call i returns logits whose arg-max (with probability 1) is script[i]"""

        def __init__(self, script, vocab=96):
            super().__init__()
            self.script, self.vocab, self.calls = script, vocab, 0

        def forward(self, idx, max_seq_length, input_pos):
            lg = torch.zeros(1, idx.size(1), self.vocab)
            lg[0, -1, self.script[self.calls]] = 1e4
            self.calls += 1
            return lg

    gold["chat"] = []
    cases = [
        ([20, 21, 22, 0, 23], ([0],), 12),                       # single-token stop
        ([20, 4, 6, 21, 4, 6, 5, 30], ([0], [4, 6, 5], [4, 7, 5]), 12),   # 3-token stop with a false start, held-back prefix
        ([20, 21, 22, 23, 24, 25, 26], ([0], [8, 9]), 6),        # budget runs out: the held-back token is never yielded
        ([8, 9, 20], ([0], [8, 9]), 10),                         # stop sequence right at the start
        ([20, 21, 22], (), 5),                                   # no stop tokens at all
        ([30, 0], ([0], [4, 6, 5]), 9),                          # short stop hits while the long buffer is not full yet
    ]
    for script, stops, max_ret in cases:
        prompt = torch.tensor([15, 16], dtype=torch.int32)
        script = script + [40] * 16
        stream = list(ref_chat.generate(Scripted(script), prompt, max_ret, max_ret, temperature=1.0, top_k=None, stop_tokens=stops))
        gold["chat"].append({"script": script, "stops": [list(s) for s in stops], "max_returned_tokens": max_ret,
                             "yields": [y.tolist() for y in stream], "ndims": [y.ndim for y in stream]})
    gold["prompt_config"] = []
    for name in ("checkpoints/stabilityai/stablelm-tuned-alpha-3b", "checkpoints/togethercomputer/RedPajama-INCITE-Chat-3B-v1",
                 "checkpoints/togethercomputer/RedPajama-INCITE-Instruct-3B-v1", "checkpoints/tiiuae/falcon-7b-instruct",
                 "checkpoints/lmsys/vicuna-7b-v1.3", "checkpoints/lmsys/longchat-7b-16k", "checkpoints/meta-llama/Llama-2-7b-chat-hf",
                 "checkpoints/stabilityai/FreeWilly2", "checkpoints/EleutherAI/pythia-70m"):
        sp, stops = ref_chat.prompt_config(Path(name), t)
        gold["prompt_config"].append({"name": name, "system_prompt": sp, "stop_tokens": [list(x) for x in stops]})
    with open(os.path.join(OUT, "cli_golden.json"), "w") as fp:
        json.dump(gold, fp, indent=1)
    print(f"[golden] cli: tokenizer files in {d.name}, {len(gold['chat'])} chat streams, {len(gold['prompt_config'])} prompt configs")


FINETUNED = {
    # name: (kind, TINY base config, extra Config kwargs)
    "adapter_neox": ("adapter", "neox", dict(adapter_prompt_length=10, adapter_start_layer=1)),
    "adapter_llama_gqa": ("adapter", "llama_gqa", dict(adapter_prompt_length=10, adapter_start_layer=1)),
    "adapter_v2_llama_mha": ("adapter_v2", "llama_mha", dict(adapter_prompt_length=10, adapter_start_layer=1)),
    "adapter_v2_falcon_mqa": ("adapter_v2", "falcon_mqa", dict(adapter_prompt_length=6, adapter_start_layer=0)),
    "lora_llama_gqa": ("lora", "llama_gqa", dict(r=4, alpha=8, to_query=True, to_key=False, to_value=True, to_projection=True,
                                                 to_mlp=True, to_head=True)),
    "lora_neox": ("lora", "neox", dict(r=2, alpha=4, to_query=True, to_key=True, to_value=True, to_projection=False, to_mlp=False,
                                       to_head=False)),
}


@torch.no_grad()
def finetuned_cases():
    """LLaMA-Adapter, Adapter v2 and merged-LoRA inference of the UNMODIFIED reference (lit_gpt/adapter.py,
    adapter_v2.py, lora.py) on tiny seeded models -> tests/golden/ft_*.npz; the oracle must reproduce them."""
    import lit_gpt.adapter as ref_adapter
    import lit_gpt.adapter_v2 as ref_adapter_v2
    import lit_gpt.lora as ref_lora

    for name, (kind, base, extra) in FINETUNED.items():
        kw = dict(TINY[base])
        seed = 1234
        ocfg = oracle_cfg(kw)
        sd = oracle.random_state_dict(ocfg, seed=seed, perturb_norm=True)
        if kind == "lora":
            cfg = ref_lora.Config(**kw, **extra)
            model = ref_lora.GPT(cfg)
            qkv_en = (extra["to_query"], extra["to_key"], extra["to_value"])
            sd.update(oracle.lora_extra_state(ocfg, seed + 1, extra["r"], qkv_en, extra["to_projection"], extra["to_mlp"], extra["to_head"]))
        else:
            cfg = ref_adapter.Config(**kw, **extra)
            model = ref_adapter.GPT(cfg)
            if kind == "adapter_v2":
                ref_adapter_v2.add_adapter_v2_parameters_to_linear_layers(model)
            sd.update(oracle.adapter_extra_state(ocfg, seed + 1, extra["adapter_start_layer"], extra["adapter_prompt_length"],
                                                 kind == "adapter_v2"))
        res = model.load_state_dict(sd, strict=True)
        assert not res.missing_keys and not res.unexpected_keys, res
        model.eval()
        osd = sd
        if kind == "lora":
            ref_lora.merge_lora_weights(model)
            osd = oracle.lora_merge_state_dict(ocfg, sd, extra["r"], extra["alpha"], qkv_en)
            merged = {k: v for k, v in model.state_dict().items() if ".lora_" not in k}
            assert set(merged) == set(osd)
            for k in merged:
                check(f"{name}/merged/{k}", osd[k], merged[k], 0.0)
        g = torch.Generator().manual_seed(1)
        B, T, steps, max_seq = 2, 7, 8, 32
        idx = torch.randint(0, cfg.padded_vocab_size, (B, T), generator=g)
        ref_full = model(idx)
        model.reset_cache()
        pos = torch.arange(T)
        ref_prefill = model(idx, max_seq, pos)
        forced = torch.randint(0, cfg.padded_vocab_size, (steps, B, 1), generator=g)
        ref_steps = []
        for s_ in range(steps):
            pos = pos[-1:] + 1
            ref_steps.append(model(forced[s_], max_seq, pos))
        ref_steps = torch.stack(ref_steps)
        model.reset_cache()
        prompt = torch.randint(0, cfg.padded_vocab_size, (5,), generator=g, dtype=torch.int64).to(torch.int32)
        ref_gen = ref_generate.generate(model, prompt, 30, 30, temperature=1.0, top_k=1)
        model.reset_cache()

        om = oracle.OracleGPT(ocfg, osd)
        d1 = check(name + "/full", om(idx), ref_full, 2e-6)
        om.reset_cache()
        pos = torch.arange(T)
        d2 = check(name + "/prefill", om(idx, max_seq, pos), ref_prefill, 2e-6)
        for s_ in range(steps):
            pos = pos[-1:] + 1
            check(name + f"/step{s_}", om(forced[s_], max_seq, pos), ref_steps[s_], 2e-6)
        om.reset_cache()
        og = oracle.generate(om, prompt, 30, 30, temperature=1.0, top_k=1, argmax_ties=True)
        assert torch.equal(og, ref_gen), (name, og, ref_gen)
        # the adapter must matter: the same model without it decodes something else
        if kind != "lora":
            plain = oracle.OracleGPT(ocfg, {k: v for k, v in sd.items() if "adapter" not in k and "gating" not in k})
            assert (plain(idx) - ref_full).abs().max() > 1e-3
        print(f"[golden] {name}: oracle == reference (max diff {max(d1, d2):.1e}); tokens equal")
        np.savez_compressed(
            os.path.join(OUT, f"ft_{name}.npz"), kind=kind, base=base,
            cfg_keys=np.array(list(kw.keys())), cfg_vals=np.array([repr(v) for v in kw.values()]),
            extra_keys=np.array(list(extra.keys())), extra_vals=np.array([repr(v) for v in extra.values()]),
            seed=seed, sd_sha256=sd_digest(sd), idx=idx.numpy(), max_seq=max_seq, forced=forced.numpy(), ref_full=ref_full.numpy(),
            ref_prefill=ref_prefill.numpy(), ref_steps=ref_steps.numpy(), prompt=prompt.numpy(), ref_gen=ref_gen.numpy(),
        )


QUANTIZER_CASES = {
    # name: (N, K, groupsize, actorder, batches of (b, T))
    # groupsize != -1 cannot be run: the reference raises at gptq.py:409-411 (`self.scales[:, j] = scale` with scale (N, 1)); its
    # own callers only use groupsize = -1 with actorder (gptq.py:499, 536, 596).  The grouped path is covered by the oracle alone.
    "perrow": (48, 256, -1, False, [(2, 16), (1, 16), (3, 16)]),
    "perrow_actorder": (40, 192, -1, True, [(2, 24), (2, 24)]),
    "oneblock": (64, 128, -1, False, [(2, 40)]),
}


@torch.no_grad()
def quantizer_cases():
    """the UNMODIFIED reference's GPTQQuantizer (quantize/gptq.py:267-431) on seeded layers and inputs ->
    tests/golden/gptq_quantizer_*.npz, and its blockwise_quantization (442-548) on a tiny model -> gptq_blockwise_*.npz; the oracle
    restatement must reproduce both."""
    import contextlib
    import io

    for name, (N, K, gs, act, shapes) in QUANTIZER_CASES.items():
        lin = torch.nn.Linear(K, N, bias=False)
        W, batches = oracle.gptq_case_inputs(N, K, shapes, seed=77)
        lin.weight.data.copy_(W)
        gq = ref_gptq.GPTQQuantizer(lin, bits=4, groupsize=gs, actorder=act)
        for x in batches:
            gq.collect_input_stats(None, (x,), None)
        H_ref = gq.H.clone()
        qmod, err = gq.quantize()
        codes = torch.empty((N, K), dtype=torch.uint8)
        codes[:, 0::2] = qmod.quant_weight & 0xF
        codes[:, 1::2] = qmod.quant_weight >> 4
        H_o, n_o = oracle.gptq_hessian(batches)
        check(name + "/H", H_o, H_ref, 0.0)
        Q_o, sc_o, ze_o, err_o, Hinv_o = oracle.gptq_quantize_layer(W, H_o, groupsize=gs, actorder=act)
        check(name + "/scales", sc_o, qmod.scales, 0.0)
        check(name + "/zeros", ze_o, qmod.zeros, 0.0)
        assert torch.equal(oracle.gptq_codes(Q_o, sc_o, ze_o, gs), codes), name
        assert abs(err_o - err) <= 1e-6 * abs(err), (err_o, err)
        print(f"[golden] gptq quantizer {name}: oracle == reference (codes, grids, Hessian); error {err:.4f}")
        np.savez_compressed(os.path.join(OUT, f"gptq_quantizer_{name}.npz"), N=N, K=K, groupsize=gs, actorder=act,
                            shapes=np.array(shapes), seed=77, W=W.numpy(), H=H_ref.numpy(), Hinv=Hinv_o.numpy(), codes=codes.numpy(),
                            scales=qmod.scales.numpy(), zeros=qmod.zeros.numpy(), error=err, nsamples=n_o)

    for base, gs in (("llama_mha", -1), ("neox", -1)):
        kw = dict(TINY[base])
        if "intermediate_size" in kw:
            kw["intermediate_size"] = 192  # in_features of every int4 layer a multiple of 32 (the kernels' K granularity)
        cfg, model, sd = build_reference(kw, 1234)
        gg = torch.Generator().manual_seed(5)
        n_samples = 24  # 1152 calibration tokens: every Hessian (K <= 256) is well conditioned
        samples = torch.randint(0, cfg.padded_vocab_size, (n_samples, cfg.block_size), generator=gg)
        with contextlib.redirect_stdout(io.StringIO()):
            ref_gptq.blockwise_quantization(model, samples, "cpu", bits=4, groupsize=gs)
        qsd = {k: v.clone() for k, v in model.state_dict().items()}
        idx = torch.randint(0, cfg.padded_vocab_size, (2, 9), generator=gg)
        model.reset_cache()
        ref_logits = model(idx)
        om = oracle.OracleGPT(oracle_cfg(kw), qsd)
        check(f"blockwise_{base}/logits", om(idx), ref_logits, 1e-5)
        plain = oracle.OracleGPT(oracle_cfg(kw), sd)(idx)
        out = {k.replace(".", "__"): v.numpy() for k, v in qsd.items() if v is not None}
        print(f"[golden] gptq blockwise {base} (groupsize {gs}): {len(qsd)} tensors; quantised vs fp32 logits max diff "
              f"{(plain - ref_logits).abs().max():.3f}")
        np.savez_compressed(os.path.join(OUT, f"gptq_blockwise_{base}.npz"), groupsize=gs, seed=1234, samples=samples.numpy(),
                            idx=idx.numpy(), ref_logits=ref_logits.numpy(), plain_logits=plain.numpy(),
                            cfg_keys=np.array(list(kw.keys())), cfg_vals=np.array([repr(v) for v in kw.values()]), **out)


if __name__ == "__main__":
    torch.manual_seed(0)
    torch.set_num_threads(8)
    if "--only-checkpoint" in sys.argv:
        checkpoint_cases()
        sys.exit(0)
    if "--only-cli" in sys.argv:
        cli_cases()
        sys.exit(0)
    if "--only-quantizer" in sys.argv:
        quantizer_cases()
        sys.exit(0)
    if "--only-finetuned" in sys.argv:
        finetuned_cases()
        sys.exit(0)
    preset_table()
    tiny_cases()
    gptq_cases()
    pythia70m()
    checkpoint_cases()
    cli_cases()
    finetuned_cases()
    quantizer_cases()
    print("[golden] all fixtures written to", OUT)
