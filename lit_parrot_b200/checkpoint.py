"""Checkpoint interop for the inference path: the files on the other side of ``GPT.load_state_dict``.

Reference behaviour restated here (not copied): ``generate/base.py:199-223`` reads ``lit_config.json`` into ``Config``,
picks ``lit_model.pth`` or — for ``quantize="gptq.int4"`` — ``lit_model_gptq.4bit.pth`` (error text of base.py:209-210),
instantiates ``GPT`` under ``quantization(quantize)`` and loads the state dict with ``strict=quantize is None`` through
``lazy_load`` (``lit_gpt/utils.py:206-220``); ``check_valid_checkpoint_dir`` (utils.py:228-262) names the missing files.

``lazy_load`` keeps the reference's contract — a context manager yielding the pickled mapping, whose tensors are not read
until they are used — without its hand-written unpickler: ``torch.load(mmap=True)`` maps the zip archive's storages, so a
tensor costs page faults only when ``load_state_dict`` copies it (one layer at a time), and 70B checkpoints never need a
second resident copy.  The values handed out are real tensors, so ``.get("model", checkpoint)``, strides (the column-major
``quant_weight`` of ``quantize/gptq.py:216-222``) and dtypes arrive exactly as saved.
"""
import json
import os
from pathlib import Path
from typing import Any, Dict, Optional, Union

import torch

from lit_parrot_b200.config import Config

PathLike = Union[str, "os.PathLike[str]"]

_GPTQ_FILE = "lit_model_gptq.4bit.pth"
_MODEL_FILE = "lit_model.pth"


class lazy_load:
    """``with lazy_load(path) as checkpoint: model.load_state_dict(checkpoint.get("model", checkpoint))``."""

    def __init__(self, fn: PathLike) -> None:
        fn = os.fspath(fn)
        if not os.path.isfile(fn):
            raise FileNotFoundError(fn)
        try:
            self.sd: Optional[Dict[str, Any]] = torch.load(fn, map_location="cpu", mmap=True, weights_only=True)
        except (RuntimeError, ValueError):
            # legacy (non-zip) files cannot be mapped; they are small by construction (pre-1.6 format)
            self.sd = torch.load(fn, map_location="cpu", weights_only=True)

    def __enter__(self) -> Dict[str, Any]:
        assert self.sd is not None, "lazy_load context already closed"
        return self.sd

    def __exit__(self, exc_type, exc_val, exc_tb) -> None:
        self.sd = None  # drops the mapping once the caller's references are gone


def check_valid_checkpoint_dir(checkpoint_dir: PathLike) -> None:
    """Raise ``SystemExit`` with a list of what is missing (the reference exits the CLI the same way, utils.py:253-262)."""
    checkpoint_dir = Path(checkpoint_dir)
    required = {
        _MODEL_FILE: (checkpoint_dir / _MODEL_FILE).is_file(),
        "lit_config.json": (checkpoint_dir / "lit_config.json").is_file(),
        "tokenizer.json OR tokenizer.model": (checkpoint_dir / "tokenizer.json").is_file()
        or (checkpoint_dir / "tokenizer.model").is_file(),
        "tokenizer_config.json": (checkpoint_dir / "tokenizer_config.json").is_file(),
    }
    if checkpoint_dir.is_dir():
        missing = [name for name, present in required.items() if not present]
        if not missing:
            return
        problem = f" is missing the files: {missing!r}"
    else:
        problem = " is not a checkpoint directory"
    local = sorted(Path("checkpoints").glob("*/*"))
    extra = ""
    if local:
        extra = "\nYou have downloaded locally:" + "".join(f"\n --checkpoint_dir {str(p.resolve())!r}" for p in local) + "\n"
    raise SystemExit(
        f"--checkpoint_dir {str(checkpoint_dir.absolute())!r}{problem}."
        "\nFind download instructions at https://github.com/Lightning-AI/lit-gpt/blob/main/tutorials\n" + extra
    )


def load_config(checkpoint_dir: PathLike) -> Config:
    """``Config(**json.load(lit_config.json))`` (base.py:201-202)."""
    with open(Path(checkpoint_dir) / "lit_config.json") as fp:
        return Config(**json.load(fp))


def checkpoint_file(checkpoint_dir: PathLike, quantize: Optional[str] = None) -> Path:
    """Model file the reference would open for this quantisation mode (base.py:206-213)."""
    checkpoint_dir = Path(checkpoint_dir)
    if quantize == "gptq.int4":
        path = checkpoint_dir / _GPTQ_FILE
        if not path.is_file():
            raise ValueError("Please run `python quantize/gptq.py` first")
        return path
    return checkpoint_dir / _MODEL_FILE


def load_checkpoint(checkpoint_dir: PathLike, quantize: Optional[str] = None, device: Union[str, torch.device, None] = "cuda",
                    dtype: Optional[torch.dtype] = None, gptq_tile_cols: int = -1, require_tokenizer: bool = False):
    """Build ``GPT`` from a lit-gpt checkpoint directory and load its weights; returns the model on ``device``.

    ``dtype``: parameter dtype of the instantiated (non-quantised) layers, the reference's ``precision`` choice; the
    checkpoint's values are cast on copy like ``load_state_dict`` does.  ``gptq_tile_cols``: group size the int4 file was
    quantised with (the scales / zeros shapes must match; -1 = per row, what the reference's context builds).
    """
    from lit_parrot_b200.model import GPT
    from lit_parrot_b200.utils import quantization

    checkpoint_dir = Path(checkpoint_dir)
    if require_tokenizer:
        check_valid_checkpoint_dir(checkpoint_dir)
    config = load_config(checkpoint_dir)
    path = checkpoint_file(checkpoint_dir, quantize)
    if not path.is_file():
        raise FileNotFoundError(str(path))
    default = torch.get_default_dtype()
    try:
        if dtype is not None:
            torch.set_default_dtype(dtype)
        qkw = {"gptq_tile_cols": gptq_tile_cols} if quantize == "gptq.int4" else {}
        with quantization(quantize, **qkw):
            model = GPT(config)
    finally:
        torch.set_default_dtype(default)
    with lazy_load(path) as checkpoint:
        model.load_state_dict(checkpoint.get("model", checkpoint), strict=quantize is None)
    model.eval()
    if device is not None:
        model = model.to(device)
    return model


def save_checkpoint(model, checkpoint_dir: PathLike, quantized: bool = False) -> Path:
    """Write ``lit_config.json`` + ``lit_model.pth`` (or the 4-bit file name) in the reference's layout — what
    ``scripts/convert_hf_checkpoint.py`` / ``quantize/gptq.py:595-596`` leave behind — for round trips and tests."""
    checkpoint_dir = Path(checkpoint_dir)
    checkpoint_dir.mkdir(parents=True, exist_ok=True)
    cfg = {k: v for k, v in vars(model.config).items() if not k.startswith("_tp")}
    with open(checkpoint_dir / "lit_config.json", "w") as fp:
        json.dump(cfg, fp)
    path = checkpoint_dir / (_GPTQ_FILE if quantized else _MODEL_FILE)
    torch.save({k: v.detach().cpu() for k, v in model.state_dict().items()}, path)
    return path
