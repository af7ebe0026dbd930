"""SURVEY Appendix B-12 deliverable: for DEEP models the north-star tolerance (logits max-abs 2e-2, cosine > 0.999, 64 identical
greedy tokens) is below the reference's OWN bf16 noise, so the claim that can be made for `precision="bf16"` is

    distance(this repo, reference fp32)  <=  distance(reference bf16-true, reference fp32)          (the reference's noise floor)

This tool measures both sides on the same B200: the UNMODIFIED reference (baseline/_ref, eager PyTorch CUDA, subprocess) on given
weights in fp32 (no TF32) and under bf16-true, and this repo's GPT on the same weights in its two modes.  Weights: seeded random
init, pre-rounded to bf16 so that every run multiplies the same stored values.

    python tests/bf16_noise_floor.py [--out profiles/r3_bf16_noise_floor.txt]

Lives under tests/ because it draws its seeded weights from the oracle (test infrastructure).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import lit_parrot_b200 as lp  # noqa: E402
from oracle import lit_oracle as O  # noqa: E402  (seeded weights only)

CASES = {
    # 24 layers each: deep enough for the bf16 rounding noise to accumulate
    "pythia-410m (NeoX: parallel residual, LayerNorm, partial rotary)": dict(lp.name_to_config["pythia-410m"]),
    "llama-style 24 x 1024 (RMSNorm, SwiGLU, GQA 16/4)": dict(block_size=2048, vocab_size=32000, padding_multiple=64, n_layer=24, n_head=16,
                                                             n_embd=1024, n_query_groups=4, rotary_percentage=1.0, parallel_residual=False,
                                                             bias=False, _norm_class="RMSNorm", norm_eps=1e-5, _mlp_class="LLaMAMLP",
                                                             intermediate_size=2816),
}


def dist(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).abs().max()), float(a @ b / (a.norm() * b.norm()))


def same_prefix(a, b):
    n = 0
    for x, y in zip(a.tolist(), b.tolist()):
        if x != y:
            break
        n += 1
    return n


def run_case(name, kw, dev, lines):
    cfg = lp.Config(**kw)
    sd = {k: v.bfloat16().float() for k, v in O.random_state_dict(cfg, seed=1234, perturb_norm=True).items()}
    g = torch.Generator().manual_seed(1)
    idx = torch.randint(0, cfg.vocab_size, (2, 96), generator=g)
    prompt = torch.randint(0, cfg.vocab_size, (16,), generator=g).to(torch.int32)
    new_tokens = 64
    with tempfile.TemporaryDirectory() as td:
        job, out = os.path.join(td, "job.pt"), os.path.join(td, "out.pt")
        torch.save({"config": kw, "state_dict": sd, "idx": idx, "prompt": prompt, "new_tokens": new_tokens}, job)
        r = subprocess.run([sys.executable, os.path.join(REPO, "baseline", "ref_runner.py"), "cuda-logits", "--job", job, "--out", out],
                           capture_output=True, text=True, timeout=1200)
        if r.returncode != 0:
            raise RuntimeError(f"reference run failed: {r.stderr[-800:]}")
        ref = torch.load(out)
    ours = {}
    for tag, dtype, prec in (("fp32act", torch.bfloat16, "fp32"), ("bf16", torch.bfloat16, "bf16")):
        m = lp.GPT(cfg)
        m.load_state_dict(sd)
        m = m.to(device=dev, dtype=dtype).eval().set_precision(prec)
        ours["logits_" + tag] = m._forward_impl(idx.to(dev), None, None).float().cpu()
        m.reset_cache()
        n = prompt.numel() + new_tokens
        ours["gen_" + tag] = lp.generate(m, prompt.to(dev), n, n, temperature=1.0, top_k=1).cpu()
        del m
        torch.cuda.empty_cache()
    res = {}
    rows = [("reference bf16-true vs reference fp32   (the noise floor)", ref["logits_bf16"], ref["logits_fp32"], ref["gen_bf16"], ref["gen_fp32"]),
            ("this repo precision='bf16' vs reference fp32", ours["logits_bf16"], ref["logits_fp32"], ours["gen_bf16"], ref["gen_fp32"]),
            ("this repo precision='bf16' vs reference bf16-true", ours["logits_bf16"], ref["logits_bf16"], ours["gen_bf16"], ref["gen_bf16"]),
            ("this repo fp32 activations (default) vs reference fp32", ours["logits_fp32act"], ref["logits_fp32"], ours["gen_fp32act"], ref["gen_fp32"])]
    lines.append(f"{name}: {cfg.n_layer} layers, n_embd {cfg.n_embd}; logits of 2 x 96 tokens (rms {ref['logits_fp32'].pow(2).mean().sqrt():.3f}), "
                 f"greedy 16 -> {16 + new_tokens} tokens")
    for what, a, b, ga, gb in rows:
        mx, cos = dist(a, b)
        same = same_prefix(ga[16:], gb[16:])
        lines.append(f"  {what:58s} max-abs {mx:9.2e}  cosine {cos:.6f}  identical greedy tokens {same}/{new_tokens}")
        res[what] = (mx, cos, same)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    lines = [f"bf16 noise floor on {torch.cuda.get_device_name(0)} (torch {torch.__version__}); tests/bf16_noise_floor.py"]
    allres = {}
    for name, kw in CASES.items():
        allres[name] = run_case(name, kw, dev, lines)
    text = "\n".join(lines)
    print(text)
    if args.out:
        with open(args.out, "w") as f:
            f.write(text + "\n")
    return allres


if __name__ == "__main__":
    main()
