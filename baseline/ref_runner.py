"""Drives the UNMODIFIED reference (baseline/_ref, installed by baseline/install_ref.sh) through its own public API and prints one
JSON line.  Run as a subprocess by bench.py (the reference's `lit_gpt` / `generate` / `quantize` packages have the same names as this
repo's drop-in packages, so they cannot share an interpreter's module table):

    python baseline/ref_runner.py cpu-generate --preset pythia-70m --prompt 16 --tokens 128 --reps 3
    python baseline/ref_runner.py cpu-decode   --preset stablelm-base-alpha-3b --ctx 2048 --steps 16
    python baseline/ref_runner.py cuda-decode  --preset stablelm-base-alpha-3b --ctx 2048 --steps 64 --warmup 8

Nothing of this repo's kernels, models or oracle is imported here: model = reference `lit_gpt.GPT`, loop = reference
`generate.base.generate` (cpu-generate) or the reference's `GPT.forward(idx, max_seq_length, input_pos)` (decode modes).
"""
import argparse
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
REPO = os.path.dirname(HERE)
sys.path = [p for p in sys.path if os.path.abspath(p or ".") not in (REPO, HERE)]
sys.path[:0] = [os.path.join(REF, "_shims"), REF]


def build(preset, device, dtype, quant=None):
    """Reference GPT with random-init N(0, 0.02) weights (values do not matter for timing: a 16 M pool is cycled)."""
    import torch

    import lit_gpt

    assert lit_gpt.__file__.startswith(REF), lit_gpt.__file__
    cfg = lit_gpt.Config.from_name(preset)
    prev = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        with torch.device("meta"):
            if quant:
                from lit_gpt.utils import quantization

                with quantization(quant):
                    model = lit_gpt.GPT(cfg)
            else:
                model = lit_gpt.GPT(cfg)
    finally:
        torch.set_default_dtype(prev)
    model = model.to_empty(device=device)
    g = torch.Generator(device=device).manual_seed(1234)
    pool = torch.empty(1 << 24, device=device, dtype=torch.float32).normal_(0, 0.02, generator=g)
    with torch.no_grad():
        for name, p in list(model.named_parameters()) + list(model.named_buffers()):
            if not p.is_floating_point():
                p.random_(0, 255)
            elif p.dim() >= 2:
                flat = p.view(-1)
                for s in range(0, flat.numel(), pool.numel()):
                    n = min(pool.numel(), flat.numel() - s)
                    flat[s:s + n].copy_(pool[:n])
            elif name.endswith("weight") or name.endswith("scales"):
                p.fill_(1.0 if name.endswith("weight") else 0.01)
            else:
                p.zero_()
    return model.eval(), cfg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["cpu-generate", "cpu-decode", "cuda-decode", "cuda-logits"])
    ap.add_argument("--job", default=None, help="cuda-logits: torch file {config: kwargs, state_dict, idx, prompt, new_tokens}")
    ap.add_argument("--out", default=None, help="cuda-logits: torch file to write the results to")
    ap.add_argument("--preset", default="pythia-70m")
    ap.add_argument("--prompt", type=int, default=16)
    ap.add_argument("--tokens", type=int, default=128)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--ctx", type=int, default=2048)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--budget", type=float, default=60.0)
    ap.add_argument("--quant", default=None)
    args = ap.parse_args()
    import torch

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    out = {"mode": args.mode, "preset": args.preset, "torch": torch.__version__, "threads": threads}
    if args.mode == "cpu-generate":
        # BASELINE configs[0]: the reference as-is — generate/base.py::generate on CPU, fp32, greedy (top_k = 1)
        import generate.base as gb

        torch.manual_seed(1234)
        import lit_gpt

        cfg = lit_gpt.Config.from_name(args.preset)
        model = lit_gpt.GPT(cfg)
        model.apply(model._init_weights)
        model.eval()
        prompt = torch.randint(0, cfg.vocab_size, (args.prompt,), generator=torch.Generator().manual_seed(1)).to(torch.int32)
        times = []
        for i in range(args.reps + 1):
            model.reset_cache()
            t0 = time.perf_counter()
            y = gb.generate(model, prompt, args.tokens, args.tokens, temperature=1.0, top_k=1)
            dt = time.perf_counter() - t0
            if i:
                times.append(dt)
        new = args.tokens - args.prompt
        out.update(tok_s=new / min(times), seconds=min(times), new_tokens=new, tokens_head=y[:24].tolist(),
                   what=f"reference generate() as-is: {args.preset} fp32 random init on CPU, {args.prompt}-token prompt -> {args.tokens} "
                        f"tokens, top_k=1, best of {args.reps} after 1 warm-up, {threads} threads")
        print(json.dumps(out))
        return
    if args.mode == "cuda-logits":
        # Parity aid: the reference on the GPU on GIVEN weights — full-forward logits and a greedy generation in
        # fp32 (matmul precision "highest": no TF32) and under bf16-true (parameters, activations and KV cache in bf16, eager kernels)
        import generate.base as gb
        import lit_gpt

        job = torch.load(args.job)
        device = torch.device("cuda", 0)
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        torch.set_float32_matmul_precision("highest")
        res = {}
        for tag, dtype in (("fp32", torch.float32), ("bf16", torch.bfloat16)):
            prev = torch.get_default_dtype()
            torch.set_default_dtype(dtype)  # the rope-cache dtype rule follows the default dtype (model.py:325-326), as under Fabric bf16-true
            try:
                model = lit_gpt.GPT(lit_gpt.Config(**job["config"]))
                model.load_state_dict(job["state_dict"], strict=True)
                model = model.to(device=device, dtype=dtype).eval()
                with torch.no_grad():
                    res["logits_" + tag] = model(job["idx"].to(device)).float().cpu()
                    model.reset_cache()
                    n = job["prompt"].numel() + int(job["new_tokens"])
                    res["gen_" + tag] = gb.generate(model, job["prompt"].to(device), n, n, temperature=1.0, top_k=1).cpu()
            finally:
                torch.set_default_dtype(prev)
            del model
            torch.cuda.empty_cache()
        torch.save(res, args.out)
        out.update(what="reference logits / greedy tokens in fp32 and bf16-true on the given weights", keys=sorted(res))
        print(json.dumps(out))
        return
    cuda = args.mode == "cuda-decode"
    device = torch.device("cuda", 0) if cuda else torch.device("cpu")
    dtype = torch.bfloat16 if cuda else torch.float32
    t_build = time.perf_counter()
    model, cfg = build(args.preset, device, dtype, args.quant)
    t_build = time.perf_counter() - t_build
    ctx = min(args.ctx, cfg.block_size)
    T0 = 16
    idx = torch.randint(0, cfg.vocab_size, (1, T0), generator=torch.Generator().manual_seed(1)).to(device)
    with torch.no_grad():
        # the reference sizes its KV cache and its attention by max_seq_length (model.py:130-144, 247): every decode step below
        # attends over the whole ctx-long cache, whatever the position — i.e. this IS the cost of decoding at a ctx-token context
        lg = model(idx, ctx, torch.arange(T0, device=device))
        tok = lg[:, -1].argmax(-1, keepdim=True)
        start = ctx - args.steps - args.warmup - 1
        pos = torch.tensor([start], device=device)
        for _ in range(args.warmup):
            tok = model(tok, ctx, pos)[:, -1].argmax(-1, keepdim=True)
            pos = pos + 1
        if cuda:
            torch.cuda.synchronize()
        n, t0 = 0, time.perf_counter()
        while n < args.steps and (time.perf_counter() - t0) < args.budget:
            tok = model(tok, ctx, pos)[:, -1].argmax(-1, keepdim=True)
            pos = pos + 1
            n += 1
        if cuda:
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    out.update(tok_s=n / dt, ms_per_step=1e3 * dt / max(n, 1), steps=n, ctx=ctx, dtype=str(dtype).split(".")[-1], build_s=t_build,
               what=f"reference GPT.forward (eager PyTorch {torch.__version__}, {'CUDA ' + torch.cuda.get_device_name(0) if cuda else 'CPU'}), "
                    f"{args.preset} {str(dtype).split('.')[-1]}{' ' + args.quant if args.quant else ''} random init, batch 1, {T0}-token prompt then {n} greedy decode "
                    f"steps at positions {start + args.warmup}..{start + args.warmup + n - 1} with max_seq_length {ctx} (the reference attends over the whole "
                    f"{ctx}-slot cache at every step){'' if cuda else f', {threads} threads'}")
    print(json.dumps(out))


if __name__ == "__main__":
    main()
