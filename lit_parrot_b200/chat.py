"""Chat front end (reference chat/base.py): a streaming ``generate`` that yields tokens as they are decoded and stops on
multi-token stop sequences (chat/base.py:20-95), token-by-token ``decode`` (98-117), the per-model prompt templates
(``prompt_config``, 202-290) and the REPL ``main`` (120-199).

For a ``lit_parrot_b200.GPT`` the tokens come from the device-resident decode loop (one captured graph replay per token, the id
is read back after each replay — streaming needs it on the host); any other model is driven by the reference's own loop.  The
stop-sequence bookkeeping is the same for both and mirrors the reference exactly, including its quirks: tokens still held back in
the look-behind buffer when the token budget runs out are never yielded, and a stop hit yields the held-back prefix as ONE tensor."""
import re
import sys
import time
from pathlib import Path
from typing import Iterable, Iterator, List, Optional, Tuple

import torch

from lit_parrot_b200.tokenizer import Tokenizer


def _stop_filter(tokens: Iterable[int], stop_tokens: Tuple[List[int], ...], device) -> Iterator[torch.Tensor]:
    """chat/base.py:47-95 on host integers: hold back as many tokens as the longest stop sequence, yield the oldest once it
    cannot be part of one."""
    stops = [list(map(int, s)) for s in stop_tokens]
    length = max((len(s) for s in stops), default=1)
    buffer = [-999] * length  # non-existing token
    yield_i = -1
    for t, tok in enumerate(tokens):
        buffer[min(t, length - 1)] = int(tok)
        for s in stops:
            if buffer[-len(s):] == s:
                if length > len(s):  # leftovers that are not part of the stop sequence
                    yield torch.tensor(buffer[:-len(s)], device=device)
                return
        if t - yield_i >= length:
            yield torch.tensor(buffer[0], device=device)
            buffer = buffer[1:] + buffer[:1]
            yield_i += 1


@torch.no_grad()
def generate(model: torch.nn.Module, idx: torch.Tensor, max_returned_tokens: int, max_seq_length: int, *, temperature: float = 1.0,
             top_k: Optional[int] = None, stop_tokens: Tuple[List[int], ...] = ()) -> Iterator[torch.Tensor]:
    """Takes a conditioning sequence (prompt) and yields the continuation token by token until a stop sequence is generated or
    ``max_returned_tokens`` is reached (chat/base.py:20-95)."""
    T = idx.size(0)
    assert max_returned_tokens > T
    from lit_parrot_b200.generate import device_token_stream, foreign_token_stream
    from lit_parrot_b200.model import GPT

    if isinstance(model, GPT):
        stream = device_token_stream(model, idx, max_returned_tokens, max_seq_length, temperature, top_k)
    else:
        stream = foreign_token_stream(model, idx, max_returned_tokens, max_seq_length, temperature, top_k)
    yield from _stop_filter(stream, stop_tokens, idx.device)


def decode(tokenizer: Tokenizer, token_stream: Iterator[torch.Tensor], out=None) -> int:
    """chat/base.py:98-117: print the reply as it arrives; SentencePiece needs the whole sequence re-decoded per token."""
    out = out or sys.stdout
    tokens_generated = 0
    if tokenizer.backend == "huggingface":
        for token in token_stream:
            print(tokenizer.decode(token), end="", flush=True, file=out)
            tokens_generated += 1
    elif tokenizer.backend == "sentencepiece":
        so_far: List[int] = []
        decoded_so_far = ""
        for token in token_stream:
            so_far.extend(token.view(-1).tolist())
            decoded_new = tokenizer.decode(torch.tensor(so_far))
            print(decoded_new[len(decoded_so_far):], end="", flush=True, file=out)
            decoded_so_far = decoded_new
            tokens_generated += 1
    else:
        raise NotImplementedError(tokenizer.backend)
    return tokens_generated


def prompt_config(checkpoint_dir: Path, tokenizer: Tokenizer) -> Tuple[str, Tuple[List[int], ...]]:
    """(system prompt template, stop sequences) per model family, selected by the checkpoint path like chat/base.py:202-290.
    The template texts are the model vendors' published chat formats (data, restated as in the reference)."""
    name = str(checkpoint_dir)
    eos = [tokenizer.eos_id]
    tid = tokenizer.token_to_id
    if re.search(r"stabilityai.*tuned-alpha", name):
        system_prompt = (
            "<|SYSTEM|># StableLM Tuned (Alpha version)\n- StableLM is a helpful and harmless open-source AI language"
            " model developed by StabilityAI.\n- StableLM is excited to be able to help the user, but will refuse to do"
            " anything that could be considered harmful to the user.\n- StableLM is more than just an information"
            " source, StableLM is also able to write poetry, short stories, and make jokes.\n- StableLM will refuse to"
            " participate in anything that could harm a human.<|USER|>{prompt}<|ASSISTANT|>")
        return system_prompt, (eos, [tid("<|SYSTEM|>")], [tid("<|ASSISTANT|>")], [tid("<|USER|>")])
    if re.search(r"togethercomputer.*Chat", name):
        lt, gt = tid("<"), tid(">:")
        return "<human>: {prompt}\n<bot>:", (eos, [lt, tid("human"), gt], [lt, tid("bot"), gt])
    if re.search(r"togethercomputer.*Instruct", name):
        colon = tid(":")
        return "Q: {prompt}\nA:", (eos, [tid("Q"), colon], [tid("Question")], [tid("A"), colon], [tid("Label"), colon],
                                    [187, 187], [535], [2756])  # '\n' '\n' | '\n\n' | '\n\n\n'
    if re.search(r"falcon.*-instruct", name):
        return ("Do not prefix your replies with 'Bot: '\nUser: {prompt}\n",
                (eos, [tid("User"), tid(":")], [193, tid("User")]))  # 193: '\n'
    if re.search(r"vicuna|longchat", name):
        return ("A chat between a curious user and an artificial intelligence assistant. The assistant gives helpful, "
                "detailed, and polite answers to the user's questions. USER: {prompt} ASSISTANT:", (eos,))
    if re.search("Llama-2.*-chat", name):
        b_inst, e_inst = "[INST]", "[/INST]"
        b_sys, e_sys = "<<SYS>>\n", "\n<</SYS>>\n\n"
        system_prompt = (
            f"{b_inst} {b_sys}You are a helpful, respectful and honest assistant. Always answer as helpfully as"
            " possible, while being safe.  Your answers should not include any harmful, unethical, racist, sexist,"
            " toxic, dangerous, or illegal content. Please ensure that your responses are socially unbiased and"
            " positive in nature.\n\nIf a question does not make any sense, or is not factually coherent, explain why"
            " instead of answering something not correct. If you don't know the answer to a question, please don't"
            f" share false information.{e_sys} {{prompt}} {e_inst} ")
        return system_prompt, (eos,)
    if re.search("FreeWilly2", name):
        return ("### System:\nThis is a system prompt, please behave and help the user.\n\n### User:\n{prompt}\n\n### Assistant:\n",
                (eos,))
    return "{prompt}", (eos,)


def main(*, top_k: int = 200, temperature: float = 0.8, checkpoint_dir: Path = Path("checkpoints/stabilityai/stablelm-tuned-alpha-3b"),
         quantize: Optional[str] = None, precision: str = "bf16-true") -> None:
    """Starts a conversation with a tuned GPT model (chat/base.py:120-199)."""
    from lit_parrot_b200.cli import load_model

    checkpoint_dir = Path(checkpoint_dir)
    if not torch.cuda.is_available():
        raise RuntimeError("lit_parrot_b200 runs on a CUDA (sm_100a) device only")
    device = torch.device("cuda", torch.cuda.current_device())
    model = load_model(checkpoint_dir, quantize, precision, device, log=print)
    tokenizer = Tokenizer(checkpoint_dir)
    system_prompt, stop_tokens = prompt_config(checkpoint_dir, tokenizer)
    while True:
        try:
            prompt = input(">> Prompt: ")
        except KeyboardInterrupt:
            break
        if not prompt:
            break
        encoded_prompt = tokenizer.encode(system_prompt.format(prompt=prompt), device=device)
        max_returned_tokens = model.config.block_size
        y = generate(model, encoded_prompt, max_returned_tokens, max_seq_length=max_returned_tokens, temperature=temperature,
                     top_k=top_k, stop_tokens=stop_tokens)
        print(">> Reply: ", end="")
        try:
            t0 = time.perf_counter()
            tokens_generated = decode(tokenizer, y)
            t = time.perf_counter() - t0
            model.reset_cache()
            print(f"\nTime for inference: {t:.02f} sec total, {tokens_generated / t:.02f} tokens/sec", file=sys.stderr)
        except KeyboardInterrupt:  # support stopping generation
            pass
        print()
