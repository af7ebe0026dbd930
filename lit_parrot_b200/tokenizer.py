"""``Tokenizer`` — same constructor, attributes and methods as the reference's (lit_gpt/tokenizer.py:8-77): a checkpoint directory
with ``tokenizer.model`` (SentencePiece; takes precedence) or ``tokenizer.json`` + ``tokenizer_config.json`` (HF ``tokenizers``).
Host-side text <-> ids only; nothing here touches the GPU path except the device the id tensor is created on."""
import json
from pathlib import Path
from typing import Optional

import torch


class Tokenizer:
    def __init__(self, checkpoint_dir: Path) -> None:
        checkpoint_dir = Path(checkpoint_dir)
        sp, hf = checkpoint_dir / "tokenizer.model", checkpoint_dir / "tokenizer.json"
        if sp.is_file():
            from sentencepiece import SentencePieceProcessor

            self.processor = SentencePieceProcessor(model_file=str(sp))
            self.backend = "sentencepiece"
            self.bos_id, self.eos_id = self.processor.bos_id(), self.processor.eos_id()
        elif hf.is_file():
            from tokenizers import Tokenizer as HFTokenizer

            self.processor = HFTokenizer.from_file(str(hf))
            self.backend = "huggingface"
            with open(checkpoint_dir / "tokenizer_config.json") as fp:
                config = json.load(fp)
            bos = config.get("bos_token")
            self.bos_id = None if bos is None else self.token_to_id(bos)
            self.eos_id = self.token_to_id(config["eos_token"])
        else:
            raise NotImplementedError

    @property
    def vocab_size(self) -> int:
        if self.backend == "huggingface":
            return self.processor.get_vocab_size(with_added_tokens=False)
        return self.processor.vocab_size()

    def token_to_id(self, token: str) -> int:
        id_ = self.processor.token_to_id(token) if self.backend == "huggingface" else self.processor.piece_to_id(token)
        if id_ is None:
            raise ValueError(f"token {token!r} not found in the collection.")
        return id_

    def encode(self, string: str, device: Optional[torch.device] = None, bos: bool = False, eos: bool = False,
               max_length: int = -1) -> torch.Tensor:
        tokens = self.processor.encode(string).ids if self.backend == "huggingface" else self.processor.encode(string)
        if bos:
            if self.bos_id is None:
                raise NotImplementedError("This tokenizer does not defined a bos token")
            tokens = [self.bos_id] + tokens
        if eos:
            tokens = tokens + [self.eos_id]
        if max_length > 0:
            tokens = tokens[:max_length]
        return torch.tensor(tokens, dtype=torch.int, device=device)

    def decode(self, tensor: torch.Tensor) -> str:
        tokens = [tensor.item()] if tensor.ndim == 0 else tensor.tolist()
        return self.processor.decode(tokens)
