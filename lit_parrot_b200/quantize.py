"""Weight-only quantised linear layers for the B200 kernels.

* ``ColBlockQuantizedLinear`` keeps the reference's constructor, buffer names, shapes and nibble order
  (quantize/gptq.py:205-252) so a ``lit_model_gptq.4bit.pth`` state dict loads unchanged; at engine-build
  time the column-major bytes are transposed once into the row-major ``LP_W_INT4`` layout the GEMV kernels stream.
* ``Linear4bit`` / ``InferenceLinear8bitLt`` stand in for the bitsandbytes classes the reference subclasses
  (quantize/bnb.py:18-75).  bitsandbytes is a third-party dependency that is not vendored by the reference;
  the quantisers below restate its published NF4 / FP4 / row-wise int8 formats (PARITY UNPINNED, see oracle header).
  int8 here is weight-only (activations stay float): closer to the unquantised model than bnb's LLM.int8().

The quantisers run with torch ops at load time (not on the hot path).
"""
from typing import Optional, Tuple

import torch

from lit_parrot_b200 import _lib

NF4_CODE = [-1.0, -0.6961928009986877, -0.5250730514526367, -0.39491748809814453, -0.28444138169288635,
            -0.18477343022823334, -0.09105003625154495, 0.0, 0.07958029955625534, 0.16093020141124725,
            0.24611230194568634, 0.33791524171829224, 0.44070982933044434, 0.5626170039176941,
            0.7229568362236023, 1.0]


def _packed_linear(*a, **k):
    from lit_parrot_b200.engine import PackedLinear

    return PackedLinear(*a, **k)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


# ------------------------------------------------------------------------------------------------
# GPTQ int4
# ------------------------------------------------------------------------------------------------
def rtn_int4_params(w: torch.Tensor, tile_cols: int, scale_dtype: Optional[torch.dtype] = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Round-to-nearest int4 on the reference's asymmetric min/max grid (quantize/gptq.py:313-347), per output row
    and per group of `tile_cols` columns.  Returns (q uint8 (N, K) values 0..15, scales (N, n_tiles), zeros).
    `scale_dtype`: the dtype the scales are STORED in (the reference keeps them in the model dtype, gptq.py:223-226 — bf16 under
    `bf16-true`); they are rounded to it before the weights are quantised against them, so q, scale, zero stay self-consistent."""
    N, K = w.shape
    if tile_cols == -1:
        tile_cols = K
    n_tiles = -(-K // tile_cols)
    wf = w.float()
    pad = n_tiles * tile_cols - K
    if pad:
        # a ragged last group only sees its real columns; replicate the last column as harmless padding
        wf = torch.cat((wf, wf[:, -1:].expand(N, pad)), dim=1)
    g = wf.view(N, n_tiles, tile_cols)
    zero_t = torch.zeros((), device=w.device)
    lo = torch.minimum(g.amin(dim=2), zero_t)
    hi = torch.maximum(g.amax(dim=2), zero_t)
    dead = (lo == 0) & (hi == 0)
    lo = torch.where(dead, torch.full_like(lo, -1.0), lo)
    hi = torch.where(dead, torch.full_like(hi, 1.0), hi)
    scales = (hi - lo) / 15
    if scale_dtype is not None:
        scales = scales.to(scale_dtype).float()
    zeros = torch.round(-lo / scales)
    q = torch.clamp(torch.round(g / scales[:, :, None]) + zeros[:, :, None], 0, 15).to(torch.uint8)
    return q.view(N, -1)[:, :K].contiguous(), scales, zeros


def tile_major_aux(scales: torch.Tensor, zeros: torch.Tensor) -> Tuple[Optional[torch.Tensor], int]:
    """Scale / zero pairs in the layout the streaming int4 kernel fetches with one bulk copy per stage:
    [N/16 tiles][n_groups][16 rows].  One 32-bit word per pair (bf16 scale bits << 16 | bf16 zero bits) when that is exact
    — scales of a bf16 checkpoint, zeros integral in [0, 255] (quantize/gptq.py:337 rounds them) — else float2.
    Returns (buffer, lp_weight.flags); (None, 0) when N is not a multiple of 16 (the exact CUDA-core kernel is used)."""
    N, ng = scales.shape
    if N % 16:
        return None, 0
    sc, ze = scales.detach().float(), zeros.detach().float()
    tile = lambda t: t.view(N // 16, 16, ng).permute(0, 2, 1).contiguous()  # noqa: E731
    exact = bool((sc.bfloat16().float() == sc).all()) and bool(((ze == ze.round()) & (ze >= 0) & (ze <= 255)).all())
    if exact:
        word = (sc.bfloat16().view(torch.int16).to(torch.int32) << 16) | (ze.bfloat16().view(torch.int16).to(torch.int32) & 0xFFFF)
        return tile(word), _lib.LP_WF_AUX_PACKED
    return torch.stack((tile(sc), tile(ze)), dim=-1).contiguous(), 0


class ColBlockQuantizedLinear(torch.nn.Module):
    def __init__(self, in_features: int, out_features: int, bias: bool, *, bits: int = 4, tile_cols: int = -1,
                 device=None, dtype=None) -> None:
        super().__init__()
        if bits != 4:
            raise NotImplementedError("the B200 kernels implement 4-bit GPTQ weights only")
        self.in_features = in_features
        self.out_features = out_features
        self.tile_cols = tile_cols if tile_cols != -1 else in_features
        self.bits = bits
        self.entries_per_byte = 2
        assert in_features % 2 == 0
        # (out, in/2) with strides (1, out): the reference's storage (gptq.py:216-222)
        self.register_buffer(
            "quant_weight", torch.empty((in_features // 2, out_features), dtype=torch.uint8, device=device).t())
        n_tiles = (in_features + self.tile_cols - 1) // self.tile_cols
        self.register_buffer("scales", torch.empty((out_features, n_tiles), device=device, dtype=dtype))
        self.register_buffer("zeros", torch.empty_like(self.scales))
        assert isinstance(bias, bool)
        self.register_buffer("bias", torch.empty((out_features,), device=device, dtype=dtype) if bias else None)
        self._lp_rows: Optional[torch.Tensor] = None  # row-major LP_W_INT4 storage once packed

    # -- reference API (offline helpers, torch ops) ------------------------------------------------
    def pack_weight(self, weight: torch.Tensor) -> None:
        """gptq.py:233-241 (float -> uint8 cast truncates, as in the reference)."""
        weight = weight.to(device=self.quant_weight.device, copy=True).float()
        for j in range(self.scales.size(1)):
            sl = slice(j * self.tile_cols, (j + 1) * self.tile_cols)
            weight[:, sl] /= self.scales[:, j:j + 1].float()
            weight[:, sl] += self.zeros[:, j:j + 1].float()
        q = weight.clamp_(min=0, max=15).to(dtype=torch.uint8)
        self.quant_weight.copy_(q[:, 0::2] | (q[:, 1::2] << 4))

    def quantize_rtn_(self, weight: torch.Tensor, scale_dtype: Optional[torch.dtype] = None) -> "ColBlockQuantizedLinear":
        """Fill the buffers from a float weight by round-to-nearest on the reference grid."""
        q, scales, zeros = rtn_int4_params(weight.to(self.quant_weight.device), self.tile_cols, scale_dtype)
        self.scales.copy_(scales)
        self.zeros.copy_(zeros)
        self.quant_weight.copy_(q[:, 0::2] | (q[:, 1::2] << 4))
        return self

    def get_weight(self, dtype: torch.dtype = torch.float) -> torch.Tensor:
        """gptq.py:243-252."""
        w = torch.empty((self.out_features, self.in_features), device=self.quant_weight.device, dtype=dtype)
        w[:, 0::2] = (self.quant_weight & 0xF).float()
        w[:, 1::2] = (self.quant_weight >> 4).float()
        for j in range(self.scales.size(1)):
            sl = slice(j * self.tile_cols, (j + 1) * self.tile_cols)
            w[:, sl] -= self.zeros[:, j:j + 1]
            w[:, sl] *= self.scales[:, j:j + 1]
        return w

    def forward(self, inp):  # pragma: no cover
        raise NotImplementedError("driven by GPT.forward through lp_linear")

    # -- kernel-side packing -----------------------------------------------------------------------
    def _rows(self) -> torch.Tensor:
        """Row-major [N, Kp/2] bytes; `quant_weight` becomes a view of it (same logical content)."""
        N, K = self.out_features, self.in_features
        lib = _lib.load()
        rb = lib.lp_int4_row_bytes(K)
        qw = self.quant_weight
        if (self._lp_rows is not None and qw.data_ptr() == self._lp_rows.data_ptr() and qw.stride(1) == 1
                and qw.stride(0) == self._lp_rows.stride(0)):
            return self._lp_rows  # already row-major (possibly interleaved with a SwiGLU partner)
        if qw.device.type != "cuda":
            raise RuntimeError("quantised weights must be on the GPU before the first forward")
        if qw.stride() != (1, N):
            qw = qw.t().contiguous().t()
        rows = torch.empty((N, rb), dtype=torch.uint8, device=qw.device)
        _lib.check(lib.lp_repack_gptq_int4(qw.data_ptr(), rows.data_ptr(), N, K, _stream()), "lp_repack_gptq_int4")
        self._lp_rows = rows
        self.quant_weight = rows[:, : K // 2]
        return rows

    def lp_pack(self):
        rows = self._rows()
        if not rows.is_contiguous():
            rows = rows.contiguous()
        bias = None if self.bias is None else self.bias.detach().float().contiguous()
        aux2, flags = tile_major_aux(self.scales, self.zeros)
        return _packed_linear(rows, _lib.LP_W_INT4, self.out_features, self.in_features, bias=bias,
                              aux0=self.scales.detach().float().contiguous(), aux1=self.zeros.detach().float().contiguous(),
                              group=self.tile_cols, aux2=aux2, flags=flags)

    def lp_pack_pair(self, other: "ColBlockQuantizedLinear"):
        """fc_1 / fc_2 interleaved row-wise for the fused SwiGLU epilogue."""
        a, b = self._rows(), other._rows()
        N, rb = a.shape
        if not (a.data_ptr() + rb == b.data_ptr() and a.stride(0) == 2 * rb):
            inter = torch.empty((2 * N, rb), dtype=torch.uint8, device=a.device)
            inter[0::2].copy_(a)
            inter[1::2].copy_(b)
            K = self.in_features
            self._lp_rows, other._lp_rows = inter[0::2], inter[1::2]
            self.quant_weight, other.quant_weight = inter[0::2][:, : K // 2], inter[1::2][:, : K // 2]
        else:
            inter = torch.as_strided(a, (2 * N, rb), (rb, 1))
        il = lambda x, y: torch.stack((x.detach().float(), y.detach().float()), dim=1).reshape(2 * N, -1).contiguous()  # noqa: E731
        bias = None if self.bias is None else il(self.bias[:, None], other.bias[:, None]).reshape(-1)
        sc, ze = il(self.scales, other.scales), il(self.zeros, other.zeros)
        aux2, flags = tile_major_aux(sc, ze)
        return _packed_linear(inter, _lib.LP_W_INT4, 2 * N, self.in_features, bias=bias, aux0=sc, aux1=ze, group=self.tile_cols,
                              aux2=aux2, flags=flags)


# ------------------------------------------------------------------------------------------------
# bitsandbytes-style 4-bit (NF4) and row-wise int8
# ------------------------------------------------------------------------------------------------
def nf4_quantize(w: torch.Tensor, blocksize: int = 64) -> Tuple[torch.Tensor, torch.Tensor]:
    """flatten -> blocks of `blocksize` -> absmax (fp32) -> nearest NF4 code; first element in the HIGH nibble."""
    code = torch.tensor(NF4_CODE, device=w.device, dtype=torch.float32)
    flat = w.detach().float().reshape(-1, blocksize)
    absmax = flat.abs().amax(dim=1)
    packed = torch.empty(flat.numel() // 2, dtype=torch.uint8, device=w.device)
    step = max(1, (1 << 22) // blocksize)
    for s in range(0, flat.shape[0], step):  # chunked: the (n, 16) distance matrix is large
        blk = flat[s:s + step] / absmax[s:s + step].clamp_min(1e-30)[:, None]
        codes = (blk.reshape(-1, 1) - code[None, :]).abs().argmin(dim=1).to(torch.uint8)
        packed[s * blocksize // 2:(s * blocksize + codes.numel()) // 2] = (codes[0::2] << 4) | codes[1::2]
    return packed, absmax


def tile_major_absmax(absmax: torch.Tensor, N: int, K: int, blocksize: int) -> Tuple[Optional[torch.Tensor], int]:
    """NF4 absmax (one fp32 per `blocksize` consecutive weights of the flattened matrix) in the layout the step kernel fetches with
    one bulk copy per 16-row x 2048-column stage: [N/16 tiles][K-stages][32 blocks][16 rows], zero padded.  (None, 0) when the
    shape is not covered (the exact CUDA-core kernel is used then)."""
    if blocksize != 64 or N % 16 or K % 256:
        return None, 0
    nblk = K // 64
    nks = -(-nblk // 32)
    a = absmax.detach().float().view(N // 16, 16, nblk)
    if nks * 32 != nblk:
        a = torch.cat((a, a.new_zeros(N // 16, 16, nks * 32 - nblk)), dim=2)
    return a.view(N // 16, 16, nks, 32).permute(0, 2, 3, 1).contiguous(), _lib.LP_WF_AUX_TILED


class Linear4bit(torch.nn.Linear):
    """NF4 weight-only linear.  Holds a float `weight` until the first forward (so float checkpoints load), then
    the weight is quantised on the GPU and replaced by the packed uint8 codes, as bitsandbytes does on `.cuda()`."""

    def __init__(self, in_features: int, out_features: int, bias: bool = True, *, quant_type: str = "nf4",
                 compress_statistics: bool = False, blocksize: int = 64, device=None, dtype=None) -> None:
        super().__init__(in_features, out_features, bias, device=device, dtype=dtype)
        if quant_type != "nf4":
            raise NotImplementedError("only quant_type='nf4' has a B200 kernel (fp4 is not implemented)")
        if compress_statistics:
            raise NotImplementedError("double quantisation of absmax (-dq) is not implemented")
        if in_features % blocksize:
            raise NotImplementedError(f"in_features must be a multiple of the NF4 block size {blocksize}")
        self.blocksize = blocksize
        # a buffer (None until the first forward quantises the weight): it then travels with .to() and state_dict()
        self.register_buffer("absmax", None)

    def forward(self, inp):  # pragma: no cover
        raise NotImplementedError("driven by GPT.forward through lp_linear")

    def _quantized(self) -> Tuple[torch.Tensor, torch.Tensor]:
        if self.weight.dtype != torch.uint8:
            if self.weight.device.type != "cuda":
                raise RuntimeError("move the model to the GPU before the first forward")
            packed, absmax = nf4_quantize(self.weight.data, self.blocksize)
            self.weight = torch.nn.Parameter(packed.view(-1, 1), requires_grad=False)
            self.absmax = absmax
        return self.weight.data.view(-1), self.absmax

    def lp_pack(self):
        packed, absmax = self._quantized()
        bias = None if self.bias is None else self.bias.detach().float().contiguous()
        aux2, flags = tile_major_absmax(absmax, self.out_features, self.in_features, self.blocksize)
        return _packed_linear(packed.contiguous(), _lib.LP_W_NF4, self.out_features, self.in_features, bias=bias, aux0=absmax,
                              group=self.blocksize, aux2=aux2, flags=flags)

    def lp_pack_pair(self, other: "Linear4bit"):
        pa, aa = self._quantized()
        pb, ab = other._quantized()
        N, K = self.out_features, self.in_features
        inter = torch.stack((pa.view(N, K // 2), pb.view(N, K // 2)), dim=1).reshape(-1).contiguous()
        am = torch.stack((aa.view(N, -1), ab.view(N, -1)), dim=1).reshape(-1).contiguous()
        bias = None
        if self.bias is not None:
            bias = torch.stack((self.bias.detach().float(), other.bias.detach().float()), dim=1).reshape(-1).contiguous()
        aux2, flags = tile_major_absmax(am, 2 * N, K, self.blocksize)
        return _packed_linear(inter, _lib.LP_W_NF4, 2 * N, K, bias=bias, aux0=am, group=self.blocksize, aux2=aux2, flags=flags)


class InferenceLinear8bitLt(torch.nn.Linear):
    """Row-wise absmax int8 weight (CB, SCB as in quantize/bnb.py:52-60), weight-only GEMV."""

    def __init__(self, in_features: int, out_features: int, bias: bool = True, device=None, dtype=None, **_unused) -> None:
        super().__init__(in_features, out_features, bias, device=device, dtype=dtype)
        self.register_buffer("SCB", None)  # row scales; a buffer so that it travels with .to() and state_dict()

    def forward(self, inp):  # pragma: no cover
        raise NotImplementedError("driven by GPT.forward through lp_linear")

    def _quantized(self) -> Tuple[torch.Tensor, torch.Tensor]:
        if self.weight.dtype != torch.int8:
            if self.weight.device.type != "cuda":
                raise RuntimeError("move the model to the GPU before the first forward")
            wf = self.weight.data.float()
            scb = wf.abs().amax(dim=1).clamp_min(1e-30)
            cb = torch.round(127.0 * wf / scb[:, None]).clamp_(-127, 127).to(torch.int8)
            self.weight = torch.nn.Parameter(cb, requires_grad=False)
            self.SCB = scb
        return self.weight.data, self.SCB

    def lp_pack(self):
        cb, scb = self._quantized()
        bias = None if self.bias is None else self.bias.detach().float().contiguous()
        return _packed_linear(cb.contiguous(), _lib.LP_W_INT8, self.out_features, self.in_features, bias=bias,
                              aux0=(scb / 127.0).contiguous())

    def lp_pack_pair(self, other: "InferenceLinear8bitLt"):
        ca, sa = self._quantized()
        cb, sb = other._quantized()
        N, K = ca.shape
        inter = torch.stack((ca, cb), dim=1).reshape(2 * N, K).contiguous()
        sc = (torch.stack((sa, sb), dim=1).reshape(-1) / 127.0).contiguous()
        bias = None
        if self.bias is not None:
            bias = torch.stack((self.bias.detach().float(), other.bias.detach().float()), dim=1).reshape(-1).contiguous()
        return _packed_linear(inter, _lib.LP_W_INT8, 2 * N, K, bias=bias, aux0=sc)
