"""ORACLE — test infrastructure, NOT product code.

A functional CPU restatement (torch ops on plain tensors, no nn.Module) of the reference's
inference hot path: ``lit_gpt.model.GPT.forward`` with KV caches and ``generate/base.py::generate``.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module, and only as the checker / CPU baseline.  The product
(``lit_parrot_b200``) never does.

Pinning: the reference is Python, so this restatement is pinned against the *unmodified* reference
imported from ``/root/reference`` in the build container (``oracle/make_golden.py`` asserts oracle == reference) and
against the golden vectors that run produced (``tests/golden/*.npz``, made by ``oracle/make_golden.py``).
The bitsandbytes NF4/int8 arithmetic is NOT in the reference tree (third-party ``bitsandbytes>=0.40.0``,
unpinned, requirements.txt:5): for those two functions the header below each says "parity unpinned".

Every function cites the reference lines it follows.  State dict keys are the reference's
(``transformer.h.{i}.attn.attn.weight`` ...), so a reference ``state_dict()`` can be fed in unchanged.
"""
import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ----------------------------------------------------------------------------------------------
# rope / norm / mlp pieces
# ----------------------------------------------------------------------------------------------
def rope_tables(block_size: int, n_elem: int, dtype: torch.dtype, condense_ratio: int = 1,
                base: int = 10000) -> Tuple[Tensor, Tensor]:
    """model.py:304-327.  theta_i = base^(-2i/n_elem); angle = (pos/condense)*theta, duplicated to
    n_elem columns; tables are cast to fp16 when the working dtype is 16-bit (model.py:325-326)."""
    theta = 1.0 / (base ** (torch.arange(0, n_elem, 2) / n_elem))
    pos = torch.arange(block_size) / condense_ratio
    ang = torch.outer(pos, theta).repeat(1, 2)
    cos, sin = torch.cos(ang), torch.sin(ang)
    if dtype in (torch.float16, torch.bfloat16, torch.int8):
        return cos.half(), sin.half()
    return cos, sin


def rotate(x: Tensor, cos: Tensor, sin: Tensor) -> Tensor:
    """model.py:330-336: rotate-half; pair is (i, i + n/2); result cast back to x's dtype."""
    half = x.size(-1) // 2
    swapped = torch.cat((-x[..., half:], x[..., :half]), dim=-1)
    return ((x * cos) + (swapped * sin)).type_as(x)


def rms_norm(x: Tensor, weight: Tensor, eps: float) -> Tensor:
    """lit_gpt/rmsnorm.py:17-21 — all arithmetic in the input dtype."""
    ms = torch.mean(x * x, dim=-1, keepdim=True)
    return weight * (x * torch.rsqrt(ms + eps))


def norm(cfg, x: Tensor, sd: Dict[str, Tensor], prefix: str) -> Tensor:
    """config.py:85-92 picks RMSNorm or torch.nn.LayerNorm (weight + bias)."""
    if cfg._norm_class == "RMSNorm":
        return rms_norm(x, sd[prefix + ".weight"], cfg.norm_eps)
    return F.layer_norm(x, (x.size(-1),), sd[prefix + ".weight"], sd[prefix + ".bias"], cfg.norm_eps)


def linear(x: Tensor, sd: Dict[str, Tensor], prefix: str) -> Tensor:
    """nn.Linear, or the GPTQ fall-back `get_weight(dtype)+F.linear` (quantize/gptq.py:263-264) when the
    state dict holds `quant_weight/scales/zeros` for this layer."""
    if prefix + ".quant_weight" in sd:
        w = gptq_dequant(sd[prefix + ".quant_weight"], sd[prefix + ".scales"], sd[prefix + ".zeros"], dtype=x.dtype)
        return F.linear(x, w, sd.get(prefix + ".bias"))
    y = F.linear(x, sd[prefix + ".weight"], sd.get(prefix + ".bias"))
    if prefix + ".adapter_scale" in sd:  # adapter v2 (lit_gpt/adapter_v2.py:34-35): adapter_scale * (linear(x) + adapter_bias)
        y = sd[prefix + ".adapter_scale"] * (y + sd[prefix + ".adapter_bias"])
    return y


def mlp(cfg, x: Tensor, sd: Dict[str, Tensor], prefix: str) -> Tensor:
    """GptNeoxMLP model.py:284-287 (exact-erf GELU) / LLaMAMLP model.py:297-301 (SwiGLU)."""
    if cfg._mlp_class == "LLaMAMLP":
        return linear(F.silu(linear(x, sd, prefix + ".fc_1")) * linear(x, sd, prefix + ".fc_2"), sd, prefix + ".proj")
    return linear(F.gelu(linear(x, sd, prefix + ".fc")), sd, prefix + ".proj")


# ----------------------------------------------------------------------------------------------
# attention with the reference's cache semantics
# ----------------------------------------------------------------------------------------------
def attention(cfg, x: Tensor, sd: Dict[str, Tensor], prefix: str, cos: Tensor, sin: Tensor, max_seq_length: int,
              mask: Optional[Tensor], input_pos: Optional[Tensor], kv: Optional[Tuple[Tensor, Tensor]],
              kv_round: Optional[torch.dtype] = None):
    """CausalSelfAttention.forward, model.py:194-254.  `kv_round`: storage precision of the cache when it is narrower than
    the activations (bf16 cache under fp32 activations): the new k / v are rounded to it before they enter the cache — the
    rounding the reference's own bf16 cache applies (model.py:236-245 with bf16 tensors)."""
    B, T, C = x.shape
    H, G, hs = cfg.n_head, cfg.n_query_groups, cfg.n_embd // cfg.n_head
    qpk = H // G
    qkv = linear(x, sd, prefix + ".attn")  # (B, T, (H+2G)*hs), rows group-interleaved [q*qpk, k, v]
    qkv = qkv.view(B, T, G, qpk + 2, hs).permute(0, 2, 3, 1, 4)  # (B, G, qpk+2, T, hs)   model.py:210-211
    q, k, v = qkv.split((qpk, 1, 1), dim=2)  # model.py:214
    if G != 1:  # model.py:217-220: MHA/GQA replicate k,v per query head; MQA keeps one head
        k = k.repeat_interleave(qpk, dim=2)
        v = v.repeat_interleave(qpk, dim=2)
    q = q.reshape(B, -1, T, hs)
    k = k.reshape(B, -1, T, hs)
    v = v.reshape(B, -1, T, hs)
    n_elem = int(cfg.rotary_percentage * hs)  # model.py:222
    q = torch.cat((rotate(q[..., :n_elem], cos, sin), q[..., n_elem:]), dim=-1)  # model.py:225-232
    k = torch.cat((rotate(k[..., :n_elem], cos, sin), k[..., n_elem:]), dim=-1)
    if kv is not None:  # model.py:234-245
        if kv_round is not None:
            k, v = k.to(kv_round).to(q.dtype), v.to(kv_round).to(q.dtype)
        ck, cv = kv
        ck, cv = ck.to(dtype=k.dtype), cv.to(dtype=v.dtype)
        if input_pos[-1] >= max_seq_length:  # sliding window by physical roll, model.py:238-242
            input_pos = torch.tensor(max_seq_length - 1)
            ck = torch.roll(ck, -1, dims=2)
            cv = torch.roll(cv, -1, dims=2)
        k = ck.index_copy_(2, input_pos, k)
        v = cv.index_copy_(2, input_pos, v)
        kv = (k, v)
    scale = 1.0 / math.sqrt(hs)  # model.py:259
    y = F.scaled_dot_product_attention(q, k, v, attn_mask=mask, dropout_p=0.0, scale=scale, is_causal=mask is None)
    if prefix + ".adapter_wte.weight" in sd:
        # LLaMA-Adapter (lit_gpt/adapter.py:234-254): attention of the same (rotated) q over the aT adaption-prompt positions,
        # whose k / v are the k / v parts of attn.attn(adapter_wte.weight) — NOT rotated, no mask — gated per head
        prefix_emb = sd[prefix + ".adapter_wte.weight"]
        aT = prefix_emb.size(0)
        aqkv = linear(prefix_emb.reshape(1, aT, C), sd, prefix + ".attn")
        aqkv = aqkv.view(1, aT, G, qpk + 2, hs).permute(0, 2, 3, 1, 4)
        _, ak, av = aqkv.split((qpk, 1, 1), dim=2)
        if G != 1:
            ak = ak.repeat_interleave(qpk, dim=2)
            av = av.repeat_interleave(qpk, dim=2)
        ak, av = ak.reshape(1, -1, aT, hs), av.reshape(1, -1, aT, hs)
        amask = torch.ones(T, aT, dtype=torch.bool)
        ay = F.scaled_dot_product_attention(q, ak, av, attn_mask=amask, dropout_p=0.0, scale=scale)
        y = y + sd[prefix + ".gating_factor"] * ay
    y = y.transpose(1, 2).contiguous().view(B, T, C)  # model.py:249
    return linear(y, sd, prefix + ".proj"), kv


def block(cfg, x, sd, i, cos, sin, max_seq_length, mask, input_pos, kv, kv_round=None):
    """Block.forward, model.py:158-180."""
    p = f"transformer.h.{i}"
    n1 = norm(cfg, x, sd, p + ".norm_1")
    h, kv = attention(cfg, n1, sd, p + ".attn", cos, sin, max_seq_length, mask, input_pos, kv, kv_round)
    if cfg.parallel_residual:
        n2 = n1 if cfg.shared_attention_norm else norm(cfg, x, sd, p + ".norm_2")
        x = x + h + mlp(cfg, n2, sd, p + ".mlp")  # (x + h) + mlp, model.py:171
    else:
        if cfg.shared_attention_norm:
            raise NotImplementedError("non-parallel residual with shared attention norm")
        x = x + h
        x = x + mlp(cfg, norm(cfg, x, sd, p + ".norm_2"), sd, p + ".mlp")
    return x, kv


class OracleGPT:
    """Holds the lazily built rope/mask/kv caches exactly like the reference module (model.py:37-39,
    79-85, 105) so it can be driven call-for-call like ``GPT.forward``."""

    def __init__(self, cfg, state_dict: Dict[str, Tensor], dtype: Optional[torch.dtype] = None,
                 kv_round: Optional[torch.dtype] = None) -> None:
        self.config = cfg
        self.kv_round = kv_round
        self.dtype = dtype or state_dict["transformer.wte.weight"].dtype
        self.sd = {k: (v.to(self.dtype) if v.is_floating_point() and not k.endswith("quant_weight") else v)
                   for k, v in state_dict.items()}
        self.rope = None
        self.mask = None
        self.kv: List[Tuple[Tensor, Tensor]] = []

    def reset_cache(self) -> None:
        self.kv.clear()

    def __call__(self, idx: Tensor, max_seq_length: Optional[int] = None, input_pos: Optional[Tensor] = None):
        """GPT.forward, model.py:63-111."""
        cfg = self.config
        B, T = idx.shape
        bs = cfg.block_size
        use_cache = input_pos is not None
        if max_seq_length is None:
            max_seq_length = bs
        if use_cache:
            assert max_seq_length >= T, f"Cannot forward sequence of length {T}, max seq length is only {max_seq_length}"
        assert max_seq_length <= bs, f"Cannot attend to {max_seq_length}, block size is only {bs}"
        assert bs >= T, f"Cannot forward sequence of length {T}, block size is only {bs}"
        hs = cfg.n_embd // cfg.n_head
        if self.rope is None:
            self.rope = rope_tables(bs, int(cfg.rotary_percentage * hs), self.dtype, cfg.condense_ratio)
        if use_cache and self.mask is None:  # dense tril bool, model.py:126-128
            self.mask = torch.tril(torch.ones((bs, bs), dtype=torch.bool))[None, None]
        cos, sin = self.rope
        if use_cache:
            cos, sin = cos.index_select(0, input_pos), sin.index_select(0, input_pos)
            mask = self.mask.index_select(2, input_pos)[:, :, :, :max_seq_length]
        else:
            cos, sin, mask = cos[:T], sin[:T], None
        x = F.embedding(idx, self.sd["transformer.wte.weight"])  # model.py:99
        if not use_cache:
            for i in range(cfg.n_layer):
                x, _ = block(cfg, x, self.sd, i, cos, sin, max_seq_length, None, None, None)
        else:
            if not self.kv:  # model.py:130-144: zero caches; MQA keeps one head, otherwise n_head heads
                heads = 1 if cfg.n_query_groups == 1 else cfg.n_head
                k_width = cos.size(-1) + hs - int(cfg.rotary_percentage * hs)  # model.py:134-139
                k_shape, v_shape = (B, heads, max_seq_length, k_width), (B, heads, max_seq_length, hs)
                self.kv = [(torch.zeros(k_shape, dtype=self.dtype), torch.zeros(v_shape, dtype=self.dtype))
                           for _ in range(cfg.n_layer)]
            for i in range(cfg.n_layer):
                x, self.kv[i] = block(cfg, x, self.sd, i, cos, sin, max_seq_length, mask, input_pos, self.kv[i], self.kv_round)
        x = norm(cfg, x, self.sd, "transformer.ln_f")
        return linear(x, self.sd, "lm_head")  # all T positions, model.py:111


# ----------------------------------------------------------------------------------------------
# generate
# ----------------------------------------------------------------------------------------------
def topk_filter(logits: Tensor, top_k: Optional[int]) -> Tensor:
    """generate/base.py:139-141: keep everything >= the k-th largest value (ties survive)."""
    if top_k is None:
        return logits
    kth = torch.topk(logits, min(top_k, logits.size(-1))).values[-1]
    return torch.where(logits < kth, torch.full_like(logits, -float("inf")), logits)


@torch.no_grad()
def generate(model, idx: Tensor, max_returned_tokens: int, max_seq_length: Optional[int] = None, *,
             temperature: float = 1.0, top_k: Optional[int] = None, eos_id: Optional[int] = None,
             argmax_ties: bool = False, logits_out: Optional[list] = None) -> Tensor:
    """generate/base.py:92-159.  ``argmax_ties=True`` replaces ``multinomial`` by lowest-index arg-max
    over the filtered logits — the deterministic reading of "greedy" (top_k=1) used for token parity
    (the reference samples uniformly among exact ties)."""
    if max_seq_length is None:
        max_seq_length = max_returned_tokens
    T = idx.size(0)
    assert max_returned_tokens > T
    out = torch.empty(max_returned_tokens, dtype=idx.dtype)
    out[:T] = idx
    input_pos = torch.arange(0, T)
    for _ in range(max_returned_tokens - T):
        x = out.index_select(0, input_pos).view(1, -1)
        logits = model(x, max_seq_length, input_pos)
        logits = logits[0, -1] / temperature
        if logits_out is not None:
            logits_out.append(logits.float().clone())
        logits = topk_filter(logits, top_k)
        if argmax_ties:
            nxt = torch.argmax(logits.float()).view(1).to(idx.dtype)
        else:
            probs = F.softmax(logits, dim=-1)
            nxt = torch.multinomial(probs, num_samples=1).to(idx.dtype)
        input_pos = input_pos[-1:] + 1
        out = out.index_copy(0, input_pos, nxt)
        if eos_id is not None and nxt.item() == eos_id:
            return out[: int(input_pos)]
    return out


# ----------------------------------------------------------------------------------------------
# weight-only quantisation formats
# ----------------------------------------------------------------------------------------------
def gptq_find_params(w: Tensor, bits: int = 4, scale_dtype: Optional[torch.dtype] = None) -> Tuple[Tensor, Tensor]:
    """Asymmetric per-row min/max grid of one column tile: quantize/gptq.py:317-347
    (perchannel=True, sym=False).  Returns (scale, zero) of shape (rows, 1); zero is an integer-valued float.
    `scale_dtype`: the dtype the scale buffer is stored in (the model dtype, gptq.py:223-226: bf16 under bf16-true); the scale is
    rounded to it BEFORE the zero point and the weights are quantised against it, so (q, scale, zero) stay self-consistent."""
    maxq = 2 ** bits - 1
    z = torch.zeros(w.shape[0])
    lo = torch.minimum(w.min(1)[0], z)
    hi = torch.maximum(w.max(1)[0], z)
    dead = (lo == 0) & (hi == 0)
    lo[dead], hi[dead] = -1, +1
    scale = (hi - lo) / maxq
    if scale_dtype is not None:
        scale = scale.to(scale_dtype).float()
    zero = torch.round(-lo / scale)
    return scale.reshape(-1, 1), zero.reshape(-1, 1)


def gptq_pack(w: Tensor, scales: Tensor, zeros: Tensor, tile_cols: int, bits: int = 4) -> Tensor:
    """quantize/gptq.py:233-241: q = clamp(w/scale + zero, 0, 15) truncated to uint8 (a float->uint8
    cast, i.e. NOT round-to-nearest); column 2j in the low nibble, 2j+1 in the high nibble; stored
    (out, in/2) with strides (1, out) (gptq.py:216-222)."""
    w = w.clone().float()
    for j in range(scales.size(1)):
        sl = slice(j * tile_cols, (j + 1) * tile_cols)
        w[:, sl] /= scales[:, j : j + 1]
        w[:, sl] += zeros[:, j : j + 1]
    q = w.clamp_(min=0, max=2 ** bits - 1).to(torch.uint8)
    per = 8 // bits
    packed = torch.zeros((w.shape[0], w.shape[1] // per), dtype=torch.uint8).t().contiguous().t()
    for nr in range(per):
        packed += q[:, nr::per] << (nr * bits)
    return packed


def gptq_rtn_quantize(w: Tensor, tile_cols: int, bits: int = 4, scale_dtype: Optional[torch.dtype] = None):
    """Round-to-nearest group quantiser built from the reference's own grid (find_params_weight) and
    its own quantise step q = clamp(round(w/scale) + zero, 0, maxq) (gptq.py:313-315), packed with the
    reference's nibble layout.  (The GPTQ error-propagation solver, gptq.py:365-445, is offline and out
    of scope; calibration data is not available offline.)"""
    if tile_cols == -1:
        tile_cols = w.shape[1]
    n_tiles = (w.shape[1] + tile_cols - 1) // tile_cols
    scales = torch.empty(w.shape[0], n_tiles)
    zeros = torch.empty(w.shape[0], n_tiles)
    maxq = 2 ** bits - 1
    q = torch.empty_like(w, dtype=torch.float32)
    for j in range(n_tiles):
        sl = slice(j * tile_cols, (j + 1) * tile_cols)
        s, z = gptq_find_params(w[:, sl].float(), bits, scale_dtype)
        scales[:, j : j + 1], zeros[:, j : j + 1] = s, z
        q[:, sl] = torch.clamp(torch.round(w[:, sl].float() / s) + z, 0, maxq)
    per = 8 // bits
    qi = q.to(torch.uint8)
    packed = torch.zeros((w.shape[0], w.shape[1] // per), dtype=torch.uint8).t().contiguous().t()
    for nr in range(per):
        packed += qi[:, nr::per] << (nr * bits)
    return packed, scales, zeros


def infer_tile_cols(in_f: int, n_tiles: int) -> int:
    """The module stores `tile_cols` (gptq.py:209); a bare state dict only has n_tiles = ceil(in/tile_cols).  Exact when
    divisible; for a ragged last tile pick the multiple of 32 that reproduces n_tiles (e.g. 4544 / 36 tiles -> 128)."""
    if in_f % n_tiles == 0:
        return in_f // n_tiles
    for t in range(32, in_f + 32, 32):
        if -(-in_f // t) == n_tiles:
            return t
    return -(-in_f // n_tiles)


def gptq_dequant(quant_weight: Tensor, scales: Tensor, zeros: Tensor, dtype=torch.float32, bits: int = 4,
                 tile_cols: Optional[int] = None) -> Tensor:
    """quantize/gptq.py:243-252: nibble -> float, then `-= zero`, `*= scale` evaluated in `dtype`
    (so in bf16 the dequantised weight is rounded to bf16 before the matmul)."""
    out_f, per = quant_weight.shape[0], 8 // bits
    in_f = quant_weight.shape[1] * per
    if tile_cols is None:
        tile_cols = infer_tile_cols(in_f, scales.shape[1])
    w = torch.empty((out_f, in_f), dtype=dtype)
    m = (1 << bits) - 1
    for nr in range(per):
        w[:, nr::per] = ((quant_weight >> (nr * bits)) & m).float()
    for j in range(scales.size(1)):
        sl = slice(j * tile_cols, (j + 1) * tile_cols)
        w[:, sl] -= zeros[:, j : j + 1]
        w[:, sl] *= scales[:, j : j + 1]
    return w


# NF4 code book of bitsandbytes (third-party, un-vendored; values re-derived from the published
# construction: normal quantiles).  PARITY UNPINNED: no reference source/tests here.
NF4_CODE = torch.tensor([
    -1.0, -0.6961928009986877, -0.5250730514526367, -0.39491748809814453, -0.28444138169288635,
    -0.18477343022823334, -0.09105003625154495, 0.0, 0.07958029955625534, 0.16093020141124725,
    0.24611230194568634, 0.33791524171829224, 0.44070982933044434, 0.5626170039176941,
    0.7229568362236023, 1.0], dtype=torch.float32)


def nf4_quantize(w: Tensor, blocksize: int = 64) -> Tuple[Tensor, Tensor]:
    """PARITY UNPINNED (bitsandbytes `quantize_4bit(quant_type="nf4")`, called through
    quantize/bnb.py:62-75 / lit_gpt/utils.py:53-60).  Published algorithm: flatten, blocks of 64,
    absmax per block (fp32), code = nearest NF4 entry of w/absmax, two codes per byte with the FIRST
    element in the HIGH nibble.  Returns (packed uint8 (n/2,), absmax fp32 (n/blocksize,))."""
    flat = w.float().reshape(-1, blocksize)
    absmax = flat.abs().amax(dim=1)
    normed = flat / absmax.clamp_min(1e-30)[:, None]
    codes = (normed.reshape(-1, 1) - NF4_CODE[None, :]).abs().argmin(dim=1).to(torch.uint8)
    packed = (codes[0::2] << 4) | codes[1::2]
    return packed, absmax


def nf4_dequantize(packed: Tensor, absmax: Tensor, shape, blocksize: int = 64, dtype=torch.float32) -> Tensor:
    """PARITY UNPINNED: w = NF4_CODE[nibble] * absmax[block] (fp32), cast to dtype."""
    codes = torch.stack(((packed >> 4) & 0xF, packed & 0xF), dim=1).reshape(-1).long()
    vals = NF4_CODE[codes].reshape(-1, blocksize) * absmax[:, None]
    return vals.reshape(shape).to(dtype)


def int8_quantize(w: Tensor) -> Tuple[Tensor, Tensor]:
    """PARITY UNPINNED (bitsandbytes row-wise `double_quant`, quantize/bnb.py:52-60): per output row
    SCB = max|w|, CB = round(127*w/SCB) as int8.  Weight-only restatement: activations stay float,
    bnb's per-token activation quantisation and >6.0 outlier split are NOT modelled (documented)."""
    wf = w.float()
    scb = wf.abs().amax(dim=1).clamp_min(1e-30)
    cb = torch.round(127.0 * wf / scb[:, None]).clamp_(-127, 127).to(torch.int8)
    return cb, scb


def int8_dequantize(cb: Tensor, scb: Tensor, dtype=torch.float32) -> Tensor:
    return (cb.float() * (scb[:, None] / 127.0)).to(dtype)


# ----------------------------------------------------------------------------------------------
# synthetic weights (shared by tests and bench so oracle and product see identical tensors)
# ----------------------------------------------------------------------------------------------
def state_dict_shapes(cfg) -> Dict[str, Tuple[int, ...]]:
    """Key -> shape of the reference ``GPT(config).state_dict()`` (model.py:24-36, 147-156, 183-192, 278-295)."""
    E, V, I = cfg.n_embd, cfg.padded_vocab_size, cfg.intermediate_size
    hs = E // cfg.n_head
    ln = cfg._norm_class == "LayerNorm"
    out: Dict[str, Tuple[int, ...]] = {"lm_head.weight": (V, E), "transformer.wte.weight": (V, E)}

    def add_norm(p):
        out[p + ".weight"] = (E,)
        if ln:
            out[p + ".bias"] = (E,)

    def add_linear(p, o, i):
        out[p + ".weight"] = (o, i)
        if cfg.bias:
            out[p + ".bias"] = (o,)

    for l in range(cfg.n_layer):
        p = f"transformer.h.{l}"
        add_norm(p + ".norm_1")
        add_linear(p + ".attn.attn", (cfg.n_head + 2 * cfg.n_query_groups) * hs, E)
        add_linear(p + ".attn.proj", E, E)
        if not cfg.shared_attention_norm:
            add_norm(p + ".norm_2")
        if cfg._mlp_class == "LLaMAMLP":
            add_linear(p + ".mlp.fc_1", I, E)
            add_linear(p + ".mlp.fc_2", I, E)
        else:
            add_linear(p + ".mlp.fc", I, E)
        add_linear(p + ".mlp.proj", E, I)
    add_norm("transformer.ln_f")
    return out


def random_state_dict(cfg, seed: int = 1234, dtype=torch.float32, perturb_norm: bool = False) -> Dict[str, Tensor]:
    """Random-init weights following ``GPT._init_weights`` (model.py:41-54): Linear/Embedding N(0, 0.02),
    biases 0, norm weight 1.  With ``perturb_norm`` the norm affine params and linear biases are drawn
    randomly as well so that tests exercise them."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    for k, shp in state_dict_shapes(cfg).items():
        is_norm = ".norm_" in k or ".ln_f." in k
        if len(shp) == 2:
            t = torch.randn(shp, generator=g) * 0.02
        elif is_norm and k.endswith(".weight"):
            t = 1.0 + (0.1 * torch.randn(shp, generator=g) if perturb_norm else torch.zeros(shp))
        else:
            t = 0.02 * torch.randn(shp, generator=g) if perturb_norm else torch.zeros(shp)
        sd[k] = t.to(dtype)
    return sd


# ----------------------------------------------------------------------------------------------
# merged-LoRA weights (lit_gpt/lora.py)
# ----------------------------------------------------------------------------------------------
def lora_merge_state_dict(cfg, sd: Dict[str, Tensor], r: int, alpha: float, enable_qkv: Tuple[bool, bool, bool]) -> Dict[str, Tensor]:
    """`merge_lora_weights` (lora.py:676-680) on a state dict: every `X.lora_A` / `X.lora_B` pair is folded into `X.weight` and
    dropped.  Plain layers (lora.py:154-164): W += (B @ A) * alpha / r.  The fused QKV layer (lora.py:338-361):
    delta = conv1d(A[None], B[..., None], groups = #enabled) * alpha / r — lora_B's rows cut into #enabled EQUAL blocks, block j
    times rows [j r, (j + 1) r) of lora_A — scattered to the rows `lora_ind` = [q rows 0..E) + [k rows E..E+kv) + [v rows E+kv..)
    of the enabled parts (lora.py:275-294), zero elsewhere (zero_pad, lora.py:296-336)."""
    out = {k: v.clone() for k, v in sd.items() if ".lora_" not in k}
    scaling = alpha / r
    E = cfg.n_embd
    kv = E // (cfg.n_head // cfg.n_query_groups)
    for key in sd:
        if not key.endswith(".lora_A"):
            continue
        base = key[: -len(".lora_A")]
        A, Bm, W = sd[key], sd[base + ".lora_B"], out[base + ".weight"]
        if base.endswith(".attn.attn"):
            ng = sum(enable_qkv)
            delta = F.conv1d(A.unsqueeze(0), Bm.unsqueeze(-1), groups=ng).squeeze(0) * scaling
            spans = [(0, E)] * enable_qkv[0] + [(E, E + kv)] * enable_qkv[1] + [(E + kv, W.size(0))] * enable_qkv[2]
            ind = torch.cat([torch.arange(a, b) for a, b in spans])
            pad = torch.zeros_like(W)
            pad.index_copy_(0, ind, delta.to(W.dtype))
            out[base + ".weight"] = W + pad
        else:
            out[base + ".weight"] = W + (Bm @ A) * scaling
    return out


def adapter_extra_state(cfg, seed: int, start_layer: int, aT: int, v2: bool) -> Dict[str, Tensor]:
    """Seeded adapter parameters for tests (the reference initialises the gate to zero and the v2 affine to identity, which would
    test nothing): `adapter_wte.weight` N(0, 0.5), `gating_factor` N(0, 1) for layers >= start_layer (adapter.py:171-177); with
    `v2` an `adapter_bias` N(0, 0.1) / `adapter_scale` 1 + N(0, 0.3) per linear layer incl. lm_head (adapter_v2.py:38-52)."""
    g = torch.Generator().manual_seed(seed)
    out: Dict[str, Tensor] = {}
    for l in range(cfg.n_layer):
        if l >= start_layer:
            out[f"transformer.h.{l}.attn.adapter_wte.weight"] = torch.randn((aT, cfg.n_embd), generator=g) * 0.5
            out[f"transformer.h.{l}.attn.gating_factor"] = torch.randn((1, cfg.n_head, 1, 1), generator=g)
    if v2:
        for k, shp in state_dict_shapes(cfg).items():
            if len(shp) == 2 and k.endswith(".weight") and "wte" not in k:
                base = k[: -len(".weight")]
                out[base + ".adapter_bias"] = torch.randn(shp[0], generator=g) * 0.1
                out[base + ".adapter_scale"] = 1.0 + torch.randn(shp[0], generator=g) * 0.3
    return out


def lora_extra_state(cfg, seed: int, r: int, enable_qkv: Tuple[bool, bool, bool], to_projection: bool, to_mlp: bool,
                     to_head: bool) -> Dict[str, Tensor]:
    """Seeded `lora_A` (r [* #enabled], in) / `lora_B` (rows, r) ~ N(0, 0.05) for the layers that carry LoRA (lora.py:479-673)."""
    g = torch.Generator().manual_seed(seed)
    out: Dict[str, Tensor] = {}
    E = cfg.n_embd
    kv = E // (cfg.n_head // cfg.n_query_groups)

    def pair(base, rows, cols, ra):
        out[base + ".lora_A"] = torch.randn((ra, cols), generator=g) * 0.05
        out[base + ".lora_B"] = torch.randn((rows, r), generator=g) * 0.05

    shapes = state_dict_shapes(cfg)
    for l in range(cfg.n_layer):
        p = f"transformer.h.{l}"
        if any(enable_qkv):
            pair(p + ".attn.attn", E * enable_qkv[0] + kv * enable_qkv[1] + kv * enable_qkv[2], E, r * sum(enable_qkv))
        if to_projection:
            pair(p + ".attn.proj", E, E, r)
        if to_mlp:
            for name in (("fc_1", "fc_2", "proj") if cfg._mlp_class == "LLaMAMLP" else ("fc", "proj")):
                o, i = shapes[f"{p}.mlp.{name}.weight"]
                pair(f"{p}.mlp.{name}", o, i, r)
    if to_head:
        pair("lm_head", cfg.padded_vocab_size, E, r)
    return out


# ----------------------------------------------------------------------------------------------
# GPTQ quantiser (quantize/gptq.py:267-431), CPU restatement
# ----------------------------------------------------------------------------------------------
def gptq_hessian(batches: List[Tensor]) -> Tuple[Tensor, int]:
    """collect_input_stats (gptq.py:349-363) over a list of input batches [b, T, K] (or [T, K]): running
    H = H * n / (n + b) + (sqrt(2 / (n + b)) X)^T (sqrt(2 / (n + b)) X) with n counted in sequences."""
    H, n = None, 0
    for inp in batches:
        if inp.dim() == 2:
            inp = inp.unsqueeze(0)
        b = inp.shape[0]
        x = inp.reshape(-1, inp.shape[-1]).t()
        if H is None:
            H = torch.zeros((x.shape[0], x.shape[0]))
        H *= n / (n + b)
        n += b
        x = math.sqrt(2 / n) * x.float()
        H += x.matmul(x.t())
    return H, n


def gptq_row_grid(x: Tensor, maxq: int, sym: bool = False) -> Tuple[Tensor, Tensor]:
    """find_params_weight (gptq.py:318-347), perchannel=True: per-row grid of x [N, cols] -> (scale, zero) [N, 1]."""
    tmp = torch.zeros(x.shape[0])
    xmin = torch.minimum(x.min(1)[0], tmp)
    xmax = torch.maximum(x.max(1)[0], tmp)
    if sym:
        xmax = torch.maximum(torch.abs(xmin), xmax)
        neg = xmin < 0
        xmin[neg] = -xmax[neg]
    both = (xmin == 0) & (xmax == 0)
    xmin[both] = -1
    xmax[both] = +1
    scale = (xmax - xmin) / maxq
    zero = torch.full_like(scale, (maxq + 1) / 2) if sym else torch.round(-xmin / scale)
    return scale.reshape(-1, 1), zero.reshape(-1, 1)


def gptq_quantize_layer(weight: Tensor, H: Tensor, bits: int = 4, blocksize: int = 128, percdamp: float = 0.01, groupsize: int = -1,
                        actorder: bool = False, sym: bool = False):
    """GPTQQuantizer.quantize (gptq.py:365-431) on a weight [N, K] and its Hessian [K, K].  Returns the de-quantised weights Q,
    the per-group (scales, zeros) [N, n_groups], the summed loss and the upper Cholesky factor of the inverse Hessian (in the
    permuted column order when actorder)."""
    maxq = 2 ** bits - 1
    W = weight.detach().to(dtype=torch.float, copy=True)
    N, K = W.shape
    tile = K if groupsize == -1 else groupsize
    scales = torch.zeros((N, (K + tile - 1) // tile))
    zeros = torch.zeros_like(scales)
    scale, zero = gptq_row_grid(W, maxq, sym)  # gptq.py:368-370
    scales[:] = scale
    zeros[:] = zero
    H = H.clone()
    dead = torch.diag(H) == 0
    H[dead, dead] = 1
    W[:, dead] = 0
    if actorder:  # gptq.py:378-381
        perm = torch.argsort(torch.diag(H), descending=True)
        W = W[:, perm]
        H = H[perm][:, perm]
    Losses = torch.zeros_like(W)
    Q = torch.zeros_like(W)
    damp = percdamp * torch.mean(torch.diag(H))
    diag = torch.arange(K)
    H[diag, diag] += damp
    H = torch.linalg.cholesky(H)
    H = torch.cholesky_inverse(H)
    Hinv = torch.linalg.cholesky(H, upper=True)
    for i1 in range(0, K, blocksize):  # gptq.py:393-424
        i2 = min(i1 + blocksize, K)
        W1 = W[:, i1:i2].clone()
        Q1 = torch.zeros_like(W1)
        Err1 = torch.zeros_like(W1)
        Losses1 = torch.zeros_like(W1)
        Hinv1 = Hinv[i1:i2, i1:i2]
        for i in range(i2 - i1):
            w = W1[:, i]
            d = Hinv1[i, i]
            if groupsize != -1 and (i1 + i) % groupsize == 0:  # the group's grid from the CURRENT global W (not W1)
                scale, zero = gptq_row_grid(W[:, (i1 + i):(i1 + i + groupsize)], maxq, sym)
                scales[:, (i1 + i) // groupsize] = scale.squeeze(1)
                zeros[:, (i1 + i) // groupsize] = zero.squeeze(1)
            q = torch.clamp(torch.round(w.unsqueeze(1) / scale) + zero, 0, maxq)  # quantize_weight, gptq.py:313-316
            q = (scale * (q - zero)).squeeze(1)
            Q1[:, i] = q
            Losses1[:, i] = (w - q) ** 2 / d ** 2
            err1 = (w - q) / d
            W1[:, i:] -= err1.unsqueeze(1).matmul(Hinv1[i, i:].unsqueeze(0))
            Err1[:, i] = err1
        Q[:, i1:i2] = Q1
        Losses[:, i1:i2] = Losses1 / 2
        W[:, i2:] -= Err1.matmul(Hinv[i1:i2, i2:])
    if actorder:
        Q = Q[:, torch.argsort(perm)]
    return Q, scales, zeros, torch.sum(Losses).item(), Hinv


def gptq_codes(Q: Tensor, scales: Tensor, zeros: Tensor, tile_cols: int, bits: int = 4) -> Tensor:
    """The integer codes pack_weight stores (gptq.py:233-241): trunc(clamp(Q / scale + zero)), per column group."""
    K = Q.shape[1]
    cols = torch.arange(K) // (K if tile_cols == -1 else tile_cols)
    return (Q / scales[:, cols] + zeros[:, cols]).clamp_(0, 2 ** bits - 1).to(torch.uint8)


def gptq_case_inputs(N: int, K: int, shapes, seed: int = 77) -> Tuple[Tensor, List[Tensor]]:
    """Seeded layer weight [N, K] ~ N(0, 0.05) and correlated calibration batches [(b, T, K)] (a random mixing matrix, so the
    Hessian is far from diagonal) shared by oracle/make_golden.py::quantizer_cases and the tests."""
    g = torch.Generator().manual_seed(seed)
    W = torch.randn((N, K), generator=g) * 0.05
    mix = torch.randn((K, K), generator=g) / math.sqrt(K) + torch.eye(K)
    return W, [torch.randn((int(b), int(T), K), generator=g) @ mix for b, T in shapes]
