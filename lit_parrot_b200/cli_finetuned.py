"""The reference's instruction-tuned entry points: ``python generate/adapter.py``, ``generate/adapter_v2.py``
(reference generate/adapter.py:23-129, generate/adapter_v2.py:25-132), ``generate/lora.py`` (generate/lora.py:28-146) and ``generate/full.py`` (generate/full.py:23-117).  Same
arguments, same checkpoint files (base ``lit_model.pth`` + the fine-tuned adapter / LoRA file, merged into one state dict), same
Alpaca prompt (scripts/prepare_alpaca.py:141-155), same output lines.  One process drives one B200 (see cli.py)."""
import json
import sys
import time
from pathlib import Path
from typing import Optional

import torch

from lit_parrot_b200 import adapter as _adapter
from lit_parrot_b200 import adapter_v2 as _adapter_v2
from lit_parrot_b200 import lora as _lora
from lit_parrot_b200.checkpoint import check_valid_checkpoint_dir, lazy_load
from lit_parrot_b200.cli import _QUANT, _param_dtype
from lit_parrot_b200.config import Config
from lit_parrot_b200.model import GPT as BaseGPT
from lit_parrot_b200.generate import generate
from lit_parrot_b200.tokenizer import Tokenizer
from lit_parrot_b200.utils import quantization


def generate_prompt(example: dict) -> str:
    """Alpaca prompt with / without an input field (scripts/prepare_alpaca.py:141-155)."""
    head = "Below is an instruction that describes a task"
    tail = "Write a response that appropriately completes the request.\n\n"
    if example["input"]:
        return (f"{head}, paired with an input that provides further context. {tail}"
                f"### Instruction:\n{example['instruction']}\n\n### Input:\n{example['input']}\n\n### Response:")
    return f"{head}. {tail}### Instruction:\n{example['instruction']}\n\n### Response:"


def _run(kind: str, prompt: str, input: str, extra_path: Path, checkpoint_dir: Path, quantize: Optional[str], max_new_tokens: int,
         top_k: int, temperature: float, strategy: str, devices: int, precision: str, lora_kwargs: Optional[dict] = None) -> str:
    checkpoint_dir, extra_path = Path(checkpoint_dir), Path(extra_path)
    if quantize is not None and quantize not in _QUANT:
        raise ValueError(f"unknown quantize mode {quantize!r}")
    if devices > 1 or strategy == "fsdp":
        raise NotImplementedError("one process drives one GPU here; the reference's FSDP path is not reproduced")
    if not torch.cuda.is_available():
        raise RuntimeError("lit_parrot_b200 runs on a CUDA (sm_100a) device only")
    device = torch.device("cuda", torch.cuda.current_device())
    check_valid_checkpoint_dir(checkpoint_dir)
    with open(checkpoint_dir / "lit_config.json") as fp:
        cfg_json = json.load(fp)
    if kind == "lora":
        config, model_cls = _lora.Config(**lora_kwargs, **cfg_json), _lora.GPT
    elif kind == "full":
        if quantize is not None:
            raise NotImplementedError  # generate/full.py:68-70: quantised fully fine-tuned checkpoints are not supported upstream either
        config, model_cls = Config(**cfg_json), BaseGPT
    else:
        config, model_cls = _adapter.Config(**cfg_json), _adapter.GPT
    model_file = "lit_model_gptq.4bit.pth" if quantize == "gptq.int4" else "lit_model.pth"
    if quantize == "gptq.int4" and not (checkpoint_dir / model_file).is_file():
        raise ValueError("Please run `python quantize/gptq.py` first")
    checkpoint_path = extra_path if kind == "full" else checkpoint_dir / model_file  # full: the fine-tuned file holds every weight
    print(f"Loading model {str(checkpoint_path)!r} with {config.__dict__}", file=sys.stderr)
    t0 = time.time()
    prev = torch.get_default_dtype()
    torch.set_default_dtype(_param_dtype(precision))
    try:
        with quantization(quantize):
            model = model_cls(config)
            if kind == "adapter_v2":
                _adapter_v2.add_adapter_v2_parameters_to_linear_layers(model)
    finally:
        torch.set_default_dtype(prev)
    print(f"Time to instantiate model: {time.time() - t0:.02f} seconds.", file=sys.stderr)
    t0 = time.time()
    with lazy_load(checkpoint_path) as checkpoint, lazy_load(extra_path) as extra:
        sd = dict(checkpoint.get("model", checkpoint))
        if kind != "full":
            sd.update(extra.get("model", extra))
        model.load_state_dict(sd, strict=quantize is None)
    print(f"Time to load the model weights: {time.time() - t0:.02f} seconds.", file=sys.stderr)
    model = model.eval().to(device)
    if kind == "lora":
        _lora.merge_lora_weights(model)
    tokenizer = Tokenizer(checkpoint_dir)
    encoded = tokenizer.encode(generate_prompt({"instruction": prompt, "input": input}), device=device)
    prompt_length = encoded.size(0)
    max_returned_tokens = prompt_length + max_new_tokens
    t0 = time.perf_counter()
    y = generate(model, encoded, max_returned_tokens, max_seq_length=max_returned_tokens, temperature=temperature, top_k=top_k,
                 eos_id=tokenizer.eos_id)
    t = time.perf_counter() - t0
    model.reset_cache()
    output = tokenizer.decode(y)
    output = output.split("### Response:")[1].strip()
    print(output)
    tokens_generated = y.size(0) - prompt_length
    print(f"\n\nTime for inference: {t:.02f} sec total, {tokens_generated / t:.02f} tokens/sec", file=sys.stderr)
    print(f"Memory used: {torch.cuda.max_memory_allocated() / 1e9:.02f} GB", file=sys.stderr)
    return output


def main_adapter(prompt: str = "What food do lamas eat?", input: str = "",
                 adapter_path: Path = Path("out/adapter/alpaca/lit_model_adapter_finetuned.pth"),
                 checkpoint_dir: Path = Path("checkpoints/stabilityai/stablelm-base-alpha-3b"), quantize: Optional[str] = None,
                 max_new_tokens: int = 100, top_k: int = 200, temperature: float = 0.8, strategy: str = "auto", devices: int = 1,
                 precision: str = "bf16-true") -> None:
    """Generates a response based on a given instruction and an optional input (GPT-Adapter checkpoints, generate/adapter.py:23)."""
    _run("adapter", prompt, input, adapter_path, checkpoint_dir, quantize, max_new_tokens, top_k, temperature, strategy, devices, precision)


def main_adapter_v2(prompt: str = "What food do lamas eat?", input: str = "",
                    adapter_path: Path = Path("out/adapter_v2/alpaca/lit_model_adapter_finetuned.pth"),
                    checkpoint_dir: Path = Path("checkpoints/stabilityai/stablelm-base-alpha-3b"), quantize: Optional[str] = None,
                    max_new_tokens: int = 100, top_k: int = 200, temperature: float = 0.8, strategy: str = "auto", devices: int = 1,
                    precision: str = "bf16-true") -> None:
    """Generates a response based on a given instruction and an optional input (GPT-AdapterV2 checkpoints, generate/adapter_v2.py:25)."""
    _run("adapter_v2", prompt, input, adapter_path, checkpoint_dir, quantize, max_new_tokens, top_k, temperature, strategy, devices,
         precision)


def main_full(prompt: str = "What food do lamas eat?", input: str = "",
              finetuned_path: Path = Path("out/full/alpaca/lit_model_finetuned.pth"),
              checkpoint_dir: Path = Path("checkpoints/stabilityai/stablelm-base-alpha-3b"), quantize: Optional[str] = None,
              max_new_tokens: int = 100, top_k: int = 200, temperature: float = 0.8, strategy: str = "auto", devices: int = 1,
              precision: str = "bf16-true") -> None:
    """Generates a response based on a given instruction and an optional input (fully fine-tuned checkpoints, generate/full.py:23)."""
    _run("full", prompt, input, finetuned_path, checkpoint_dir, quantize, max_new_tokens, top_k, temperature, strategy, devices, precision)


# module-level LoRA hyper-parameters of generate/lora.py:19-27 (they must match the fine-tuning run)
lora_r, lora_alpha, lora_dropout = 8, 16, 0.05
lora_query, lora_key, lora_value, lora_projection, lora_mlp, lora_head = True, False, True, False, False, False


def main_lora(prompt: str = "What food do lamas eat?", input: str = "",
              lora_path: Path = Path("out/lora/alpaca/lit_model_lora_finetuned.pth"),
              checkpoint_dir: Path = Path("checkpoints/stabilityai/stablelm-base-alpha-3b"), quantize: Optional[str] = None,
              max_new_tokens: int = 100, top_k: int = 200, temperature: float = 0.8, strategy: str = "auto", devices: int = 1,
              precision: str = "bf16-true") -> None:
    """Generates a response based on a given instruction and an optional input (LoRA checkpoints, merged; generate/lora.py:28)."""
    kw = dict(r=lora_r, alpha=lora_alpha, dropout=lora_dropout, to_query=lora_query, to_key=lora_key, to_value=lora_value,
              to_projection=lora_projection, to_mlp=lora_mlp, to_head=lora_head)
    _run("lora", prompt, input, lora_path, checkpoint_dir, quantize, max_new_tokens, top_k, temperature, strategy, devices, precision, kw)
