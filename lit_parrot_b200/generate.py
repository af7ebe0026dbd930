"""``generate()`` — drop-in for generate/base.py:92-159.

For a ``lit_parrot_b200.GPT`` on a B200 the loop runs on the device: the prompt is prefilled once, then ONE
captured CUDA graph per token (embed -> n_layer blocks -> ln_f -> lm_head -> top-k/temperature sampling ->
append token, advance position) is replayed; the host does not read anything back until the end (or every few
tokens when an ``eos_id`` has to be honoured).  The reference syncs once per layer and once per token
(model.py:238, generate/base.py:156).

Any other model (e.g. the reference's own ``GPT`` on CPU) is driven by the plain loop with the same semantics;
all arithmetic is then that model's own.
"""
from typing import Dict, Optional

import torch

from lit_parrot_b200 import _lib

_EOS_POLL = 8  # replays between host-side EOS checks
_calls: Dict[int, int] = {}


@torch.no_grad()
def generate(
    model: torch.nn.Module,
    idx: torch.Tensor,
    max_returned_tokens: int,
    max_seq_length: Optional[int] = None,
    *,
    temperature: float = 1.0,
    top_k: Optional[int] = None,
    eos_id: Optional[int] = None,
) -> torch.Tensor:
    """Takes a conditioning sequence (prompt) ``idx`` of shape (T) and continues it up to ``max_returned_tokens``.

    ``max_seq_length`` may be omitted (the later upstream signature); it then defaults to ``max_returned_tokens``.
    ``top_k=1`` is greedy decoding (lowest index on exact ties; the reference draws among ties at random).
    With ``eos_id`` the result is cut *before* the EOS token, exactly as ``idx[:input_pos]`` does (base.py:156-157).
    """
    if max_seq_length is None:
        max_seq_length = max_returned_tokens
    T = idx.size(0)
    assert max_returned_tokens > T
    from lit_parrot_b200.model import GPT

    if isinstance(model, GPT):
        if idx.device.type != "cuda":
            raise RuntimeError("lit_parrot_b200.GPT runs on CUDA (sm_100a) only; there is no CPU path")
        return _generate_on_device(model, idx, max_returned_tokens, max_seq_length, temperature, top_k, eos_id)
    return _generate_foreign(model, idx, max_returned_tokens, max_seq_length, temperature, top_k, eos_id)


def sample(logits: torch.Tensor, temperature: float = 1.0, top_k: Optional[int] = None,
           step: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The sampling tail of generate() (base.py:136-144) as one kernel: logits (rows, V) or (V,) on the GPU ->
    int32 token ids (rows,).  ``step``: optional int32 device counter that keys the Philox stream."""
    if logits.device.type != "cuda":
        raise RuntimeError("lit_parrot_b200.sample runs on CUDA only")
    lg = logits.reshape(-1, logits.shape[-1])
    if lg.dtype != torch.bfloat16:  # bf16 logits (what GPT.forward returns for a bf16 checkpoint) are read as they are
        lg = lg.float()
    lg = lg.contiguous()
    lib = _lib.init(lg.device.index)
    out = torch.empty(lg.shape[0], dtype=torch.int32, device=lg.device)
    seed = int(torch.initial_seed()) & 0xFFFFFFFFFFFFFFFF
    k = 0 if top_k is None else min(int(top_k), lg.shape[1])
    fn = lib.lp_sample_bf16 if lg.dtype == torch.bfloat16 else lib.lp_sample
    _lib.check(fn(lg.data_ptr(), lg.shape[0], lg.shape[1], float(temperature), k, seed, None if step is None else step.data_ptr(),
                  out.data_ptr(), None, None, torch.cuda.current_stream(lg.device).cuda_stream), "lp_sample")
    return out


def _generate_on_device(model, idx, max_returned_tokens, max_seq_length, temperature, top_k, eos_id) -> torch.Tensor:
    if not temperature > 0:
        raise ValueError("temperature must be > 0 (use top_k=1 for greedy decoding)")
    cfg = model.config
    if max_returned_tokens > cfg.block_size:
        raise IndexError(f"max_returned_tokens {max_returned_tokens} exceeds the RoPE table (block_size {cfg.block_size})")
    device = idx.device
    T = idx.size(0)
    # prefill: the same checks and cache construction as GPT.forward, logits of the last position only
    logits = model._forward_impl(idx.view(1, -1), max_seq_length, torch.arange(0, T, device=device), last_only=True,
                                 raw_logits=True)
    eng = model._get_engine(device)
    lib = eng.lib
    st = eng.gen_state(max(cfg.block_size, max_returned_tokens))
    seq, pos, step, tok = st["seq"], st["pos"], st["step"], st["tok"]
    seq[:T].copy_(idx)
    pos.fill_(T - 1)
    seed = int(torch.initial_seed()) & 0xFFFFFFFFFFFFFFFF
    _calls[seed] = _calls.get(seed, 0) + 1
    step.fill_((_calls[seed] * 8192) % (1 << 30))
    k = 0 if top_k is None else min(int(top_k), cfg.padded_vocab_size)
    stream = torch.cuda.current_stream(device).cuda_stream
    # first new token from the prefill logits: seq[T] = token, pos = T
    _lib.check(lib.lp_sample(logits.data_ptr(), 1, cfg.padded_vocab_size, float(temperature), k, seed, step.data_ptr(),
                             tok.data_ptr(), seq.data_ptr(), pos.data_ptr(), stream), "lp_sample")
    n_more = max_returned_tokens - T - 1
    replay = eng.decode_step(model.kv_caches, float(temperature), k, seed)
    done = 0
    cut = None
    while done < n_more:
        burst = n_more - done if eos_id is None else min(_EOS_POLL, n_more - done)
        for _ in range(burst):
            replay()
        done += burst
        if eos_id is not None:
            hit = (seq[T:T + 1 + done] == eos_id).nonzero()
            if hit.numel():
                cut = T + int(hit[0])
                break
    if eos_id is not None and cut is None:
        hit = (seq[T:max_returned_tokens] == eos_id).nonzero()
        if hit.numel():
            cut = T + int(hit[0])
    out = seq[:max_returned_tokens] if cut is None else seq[:cut]
    out = out.to(idx.dtype, copy=True)
    eng.check_step_health()  # synchronises; raises if the step kernel's watchdog fired (never a silent wrong answer)
    return out


def device_token_stream(model, idx, max_returned_tokens, max_seq_length, temperature, top_k):
    """The tokens of `_generate_on_device` one at a time (host ints): the streaming form the chat front end needs
    (chat/base.py:57-95).  Same prefill, same captured decode step; the sampled id is read back after every replay."""
    if not temperature > 0:
        raise ValueError("temperature must be > 0 (use top_k=1 for greedy decoding)")
    cfg = model.config
    if max_returned_tokens > cfg.block_size:
        raise IndexError(f"max_returned_tokens {max_returned_tokens} exceeds the RoPE table (block_size {cfg.block_size})")
    device = idx.device
    T = idx.size(0)
    logits = model._forward_impl(idx.view(1, -1), max_seq_length, torch.arange(0, T, device=device), last_only=True, raw_logits=True)
    eng = model._get_engine(device)
    st = eng.gen_state(max(cfg.block_size, max_returned_tokens))
    seq, pos, step, tok = st["seq"], st["pos"], st["step"], st["tok"]
    seq[:T].copy_(idx)
    pos.fill_(T - 1)
    seed = int(torch.initial_seed()) & 0xFFFFFFFFFFFFFFFF
    _calls[seed] = _calls.get(seed, 0) + 1
    step.fill_((_calls[seed] * 8192) % (1 << 30))
    k = 0 if top_k is None else min(int(top_k), cfg.padded_vocab_size)
    _lib.check(eng.lib.lp_sample(logits.data_ptr(), 1, cfg.padded_vocab_size, float(temperature), k, seed, step.data_ptr(),
                                 tok.data_ptr(), seq.data_ptr(), pos.data_ptr(), torch.cuda.current_stream(device).cuda_stream),
               "lp_sample")
    yield int(seq[T])
    replay = eng.decode_step(model.kv_caches, float(temperature), k, seed)
    for n in range(max_returned_tokens - T - 1):
        replay()
        yield int(seq[T + 1 + n])
    eng.check_step_health()


def foreign_token_stream(model, idx, max_returned_tokens, max_seq_length, temperature, top_k):
    """chat/base.py:57-73 for models that are not ours (the model does all the arithmetic)."""
    T = idx.size(0)
    input_pos = torch.arange(0, T, device=idx.device)
    for _ in range(max_returned_tokens - T):
        logits = model(idx.view(1, -1), max_seq_length, input_pos)
        logits = logits[0, -1] / temperature
        if top_k is not None:
            v, _ = torch.topk(logits, min(top_k, logits.size(-1)))
            logits = torch.where(logits < v[[-1]], -float("Inf"), logits)
        probs = torch.nn.functional.softmax(logits, dim=-1)
        idx = torch.multinomial(probs, num_samples=1)
        input_pos = input_pos[-1:] + 1
        yield int(idx)


def _generate_foreign(model, idx, max_returned_tokens, max_seq_length, temperature, top_k, eos_id) -> torch.Tensor:
    """generate/base.py:113-159 for models that are not ours (host loop; the model does all the arithmetic)."""
    T = idx.size(0)
    device, dtype = idx.device, idx.dtype
    buf = torch.empty(max_returned_tokens, dtype=dtype, device=device)
    buf[:T] = idx
    idx = buf
    input_pos = torch.arange(0, T, device=device)
    for _ in range(max_returned_tokens - T):
        x = idx.index_select(0, input_pos).view(1, -1)
        logits = model(x, max_seq_length, input_pos)
        logits = logits[0, -1] / temperature
        if top_k is not None:
            v, _ = torch.topk(logits, min(top_k, logits.size(-1)))
            logits = torch.where(logits < v[[-1]], -float("Inf"), logits)
        probs = torch.nn.functional.softmax(logits, dim=-1)
        idx_next = torch.multinomial(probs, num_samples=1).to(dtype=dtype)
        input_pos = input_pos[-1:] + 1
        idx = idx.index_copy(0, input_pos, idx_next)
        if idx_next == eos_id:
            return idx[:input_pos]
    return idx
