"""The callers on the other side of ``generate()``: ``python generate/base.py`` (reference generate/base.py:162-257)
and ``python chat/base.py`` (reference chat/base.py:120-199).  Same arguments, same file layout of the checkpoint directory, same
lines on stdout / stderr ("Time for inference i: ... tokens/sec", "Memory used: ... GB").  What differs is below them: no Fabric —
one process drives one B200 (`devices > 1` is refused: the reference's FSDP re-gathers every layer per token and is not reproduced;
tensor parallelism is `lit_parrot_b200.tp`), and `precision` only selects the parameter dtype ("bf16-true" / "16-true" -> bf16
weights with fp32 activations, "32-true" -> fp32 weights)."""
import argparse
import inspect
import json
import sys
import time
from pathlib import Path
from typing import Optional

import torch

from lit_parrot_b200.checkpoint import check_valid_checkpoint_dir, lazy_load
from lit_parrot_b200.config import Config
from lit_parrot_b200.generate import generate
from lit_parrot_b200.model import GPT
from lit_parrot_b200.tokenizer import Tokenizer
from lit_parrot_b200.utils import quantization

_QUANT = ("bnb.nf4", "bnb.nf4-dq", "bnb.fp4", "bnb.fp4-dq", "bnb.int8", "gptq.int4")


def _param_dtype(precision: str) -> torch.dtype:
    if precision in ("bf16-true", "bf16-mixed", "16-true", "16-mixed"):
        return torch.bfloat16
    if precision in ("32-true", "32"):
        return torch.float32
    raise ValueError(f"unsupported precision {precision!r}: use 'bf16-true' or '32-true'")


def load_model(checkpoint_dir: Path, quantize: Optional[str], precision: str, device: torch.device, log=None) -> GPT:
    """generate/base.py:199-226: config json -> GPT under quantization() -> lazy_load -> load_state_dict(strict=quantize is None)."""
    check_valid_checkpoint_dir(checkpoint_dir)
    with open(checkpoint_dir / "lit_config.json") as fp:
        config = Config(**json.load(fp))
    if quantize == "gptq.int4":
        model_file = "lit_model_gptq.4bit.pth"
        if not (checkpoint_dir / model_file).is_file():
            raise ValueError("Please run `python quantize/gptq.py` first")
    else:
        model_file = "lit_model.pth"
    checkpoint_path = checkpoint_dir / model_file
    say = log or (lambda *a, **k: None)
    say(f"Loading model {str(checkpoint_path)!r} with {config.__dict__}", file=sys.stderr)
    t0 = time.time()
    prev = torch.get_default_dtype()
    torch.set_default_dtype(_param_dtype(precision))
    try:
        with torch.device("meta"), quantization(quantize):
            model = GPT(config)
    finally:
        torch.set_default_dtype(prev)
    model = model.to_empty(device="cpu")
    say(f"Time to instantiate model: {time.time() - t0:.02f} seconds.", file=sys.stderr)
    t0 = time.time()
    with lazy_load(checkpoint_path) as checkpoint:
        model.load_state_dict(checkpoint.get("model", checkpoint), strict=quantize is None)
    say(f"Time to load the model weights: {time.time() - t0:.02f} seconds.", file=sys.stderr)
    model.eval()
    return model.to(device)


def main(
    prompt: str = "Hello, my name is",
    *,
    num_samples: int = 1,
    max_new_tokens: int = 50,
    top_k: int = 200,
    temperature: float = 0.8,
    checkpoint_dir: Path = Path("checkpoints/stabilityai/stablelm-base-alpha-3b"),
    quantize: Optional[str] = None,
    strategy: str = "auto",
    devices: int = 1,
    precision: str = "bf16-true",
) -> None:
    """Generates text samples based on a pre-trained model and tokenizer (arguments as generate/base.py:162-190)."""
    checkpoint_dir = Path(checkpoint_dir)
    if quantize is not None and quantize not in _QUANT:
        raise ValueError(f"unknown quantize mode {quantize!r}")
    if devices > 1 or strategy == "fsdp":
        raise NotImplementedError("one process drives one GPU here; multi-GPU inference is tensor parallel (lit_parrot_b200.tp), "
                                  "the reference's FSDP path (generate/base.py:194-197) is not reproduced")
    if not torch.cuda.is_available():
        raise RuntimeError("lit_parrot_b200 runs on a CUDA (sm_100a) device only")
    device = torch.device("cuda", torch.cuda.current_device())
    model = load_model(checkpoint_dir, quantize, precision, device, log=print)
    tokenizer = Tokenizer(checkpoint_dir)
    encoded = tokenizer.encode(prompt, device=device)
    prompt_length = encoded.size(0)
    max_returned_tokens = prompt_length + max_new_tokens
    assert max_returned_tokens <= model.config.block_size, (max_returned_tokens, model.config.block_size)  # maximum rope cache length
    torch.manual_seed(1234)  # L.seed_everything(1234), base.py:237
    for i in range(num_samples):
        t0 = time.perf_counter()
        y = generate(model, encoded, max_returned_tokens, max_seq_length=max_returned_tokens, temperature=temperature, top_k=top_k)
        t = time.perf_counter() - t0
        model.reset_cache()
        print(tokenizer.decode(y))
        tokens_generated = y.size(0) - prompt_length
        print(f"Time for inference {i + 1}: {t:.02f} sec total, {tokens_generated / t:.02f} tokens/sec", file=sys.stderr)
    print(f"Memory used: {torch.cuda.max_memory_allocated() / 1e9:.02f} GB", file=sys.stderr)


def CLI(fn, argv=None):
    """`jsonargparse.CLI(main)` for the entry points: every parameter of the function is a `--name value` option (jsonargparse
    turns parameters with defaults into options); the first parameter may also be given positionally."""
    sig = inspect.signature(fn)
    ap = argparse.ArgumentParser(description=(fn.__doc__ or "").strip().splitlines()[0] if fn.__doc__ else None)
    first = None
    for name, p in sig.parameters.items():
        typ = Path if isinstance(p.default, Path) else {int: int, float: float, str: str}.get(type(p.default), str)
        ap.add_argument(f"--{name}", type=typ, default=p.default)
        if first is None and p.kind is not inspect.Parameter.KEYWORD_ONLY:
            first = name
            ap.add_argument(f"{name}_positional", type=typ, nargs="?", default=None, metavar=name)
    args = vars(ap.parse_args(argv))
    if first is not None:
        pos = args.pop(f"{first}_positional")
        if pos is not None:
            args[first] = pos
    return fn(**args)
