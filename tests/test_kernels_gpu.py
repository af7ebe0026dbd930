"""Per-kernel parity on a B200: every C-ABI entry point against a torch/oracle restatement of the reference op."""
import ctypes
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from lit_parrot_b200 import _lib
from lit_parrot_b200._lib import LpWeight
from oracle import lit_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def lib():
    return _lib.init(0)


def stream():
    return torch.cuda.current_stream().cuda_stream


def f32(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


def bf16r(x):
    return x.to(torch.bfloat16).float()


def run_linear(lib, x, wt, fmt, N, K, *, bias=None, aux0=None, aux1=None, group=0, epi=0, residual=None, rnd=0, path=0,
               packed_aux=None):
    M = x.shape[0]
    out = torch.full((M, N // 2 if epi == _lib.LP_EPI_SWIGLU else N), float("nan"), device=DEV)
    p = lambda a: None if a is None else a.data_ptr()  # noqa: E731
    aux2, flags = None, 0
    if fmt == _lib.LP_W_INT4:  # tile-major scale/zero buffer for the streaming kernel (exact float2 unless asked otherwise)
        from lit_parrot_b200.quantize import tile_major_aux
        aux2, flags = tile_major_aux(aux0, aux1)
        if packed_aux is not None:
            assert bool(flags & _lib.LP_WF_AUX_PACKED) == packed_aux
    rec = LpWeight(wt.data_ptr(), p(aux0), p(aux1), p(aux2), p(bias), fmt, N, K, group, flags, 0)
    _lib.check(lib.lp_set_linear_path(path))
    try:
        rc = lib.lp_linear(x.data_ptr(), M, ctypes.byref(rec), epi, p(residual), out.data_ptr(), rnd, stream())
    finally:
        lib.lp_set_linear_path(0)
    _lib.check(rc, "lp_linear")
    torch.cuda.synchronize()
    return out


def ref_epilogue(y, epi, residual):
    if epi == _lib.LP_EPI_GELU:
        return F.gelu(y)
    if epi == _lib.LP_EPI_SWIGLU:
        return F.silu(y[:, 0::2]) * y[:, 1::2]
    if epi == _lib.LP_EPI_RESIDUAL:
        return residual + y
    return y


PATHS = [1, 0]  # 1 = FMA family only, 0 = auto (streaming family where it applies)


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("wdtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("M,N,K", [(1, 64, 64), (1, 50, 136), (3, 256, 512), (8, 96, 1024), (13, 128, 256), (1, 4672, 4544), (32, 512, 1024)])
def test_linear_dense(lib, path, wdtype, M, N, K):
    w = f32(N, K, seed=1, scale=0.05).to(wdtype)
    x, b = f32(M, K, seed=2), f32(N, seed=3)
    if wdtype == torch.float32 and K % 4 or wdtype == torch.bfloat16 and K % 8:
        pytest.skip("alignment")
    fmt = _lib.LP_W_F32 if wdtype == torch.float32 else _lib.LP_W_BF16
    ref = F.linear(x.double(), w.double(), b.double())
    for epi in (0, 1, 2, 3):
        if epi == 2 and N % 2:
            continue
        res = f32(M, N, seed=4)
        got = run_linear(lib, x, w, fmt, N, K, bias=b, epi=epi, residual=res, path=path)
        want = ref_epilogue(ref, epi, res.double()).float()
        torch.testing.assert_close(got, want, rtol=2e-5, atol=2e-5 * math.sqrt(K))


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("group", [128, -1])
@pytest.mark.parametrize("M,N,K", [(1, 64, 256), (2, 48, 384), (5, 256, 1024), (1, 96, 4544), (16, 128, 512)])
def test_linear_gptq_int4(lib, path, group, M, N, K):
    """quant_weight in the reference's column-major storage -> lp_repack_gptq_int4 -> lp_linear, against the reference
    fall-back `get_weight()+F.linear` restated by the oracle (quantize/gptq.py:243-252, 263-264)."""
    w = (torch.randn(N, K, generator=torch.Generator().manual_seed(5)) * 0.02)
    packed, scales, zeros = O.gptq_rtn_quantize(w, group)
    assert packed.stride() == (1, N)
    src = torch.empty((K // 2, N), dtype=torch.uint8, device=DEV).t()
    src.copy_(packed)
    rb = lib.lp_int4_row_bytes(K)
    rows = torch.empty((N, rb), dtype=torch.uint8, device=DEV)
    _lib.check(lib.lp_repack_gptq_int4(src.data_ptr(), rows.data_ptr(), N, K, stream()))
    torch.cuda.synchronize()
    assert torch.equal(rows[:, : K // 2].cpu(), packed.contiguous())  # bit-exact transpose
    assert int(rows[:, K // 2:].sum()) == 0
    x = f32(M, K, seed=6)
    g = K if group == -1 else group
    assert O.infer_tile_cols(K, scales.shape[1]) == g
    wd = O.gptq_dequant(packed, scales, zeros)  # fp32 dequant, as the reference does for fp32 activations
    want = F.linear(x.cpu().double(), wd.double()).float().to(DEV)
    got = run_linear(lib, x, rows, _lib.LP_W_INT4, N, K, aux0=scales.to(DEV).contiguous(), aux1=zeros.to(DEV).contiguous(),
                     group=g, path=path)
    torch.testing.assert_close(got, want, rtol=2e-5, atol=3e-5 * math.sqrt(K) * 0.3)
    # bf16-faithful mode: dequantised weight and output rounded to bf16 like the reference's bf16 run
    xb = bf16r(x)
    wdb = O.gptq_dequant(packed, scales, zeros, dtype=torch.bfloat16).double()
    wantb = bf16r(F.linear(xb.cpu().double(), wdb).float()).to(DEV)
    gotb = run_linear(lib, xb, rows, _lib.LP_W_INT4, N, K, aux0=scales.to(DEV).contiguous(), aux1=zeros.to(DEV).contiguous(),
                      group=g, rnd=1, path=1)
    mism = (gotb != wantb)
    # identical except where the fp32 sum lands within rounding noise of a bf16 tie: at most 1 ulp, rarely
    assert mism.float().mean() < 0.02
    torch.testing.assert_close(gotb, wantb, rtol=2 ** -7, atol=1e-6)


@pytest.mark.parametrize("M,N,K,group", [(1, 64, 256, 128), (2, 4096, 4096, 128), (1, 22016, 4096, 128), (4, 4096, 11008, 128),
                                         (1, 256, 2048, 256), (8, 128, 4096, -1), (1, 4672, 4608, 128)])
def test_linear_gptq_int4_stream_packed_scales(lib, M, N, K, group):
    """bf16-representable scales (a bf16 checkpoint, quantize/gptq.py under bf16-true): the streaming kernel reads one
    packed 32-bit scale|zero word per (row, group) — results identical to the float2 layout and to the oracle."""
    w = (torch.randn(N, K, generator=torch.Generator().manual_seed(5)) * 0.02)
    packed, scales, zeros = O.gptq_rtn_quantize(w, group)
    scales = scales.bfloat16().float()
    src = torch.empty((K // 2, N), dtype=torch.uint8, device=DEV).t()
    src.copy_(packed)
    rows = torch.empty((N, lib.lp_int4_row_bytes(K)), dtype=torch.uint8, device=DEV)
    _lib.check(lib.lp_repack_gptq_int4(src.data_ptr(), rows.data_ptr(), N, K, stream()))
    x = f32(M, K, seed=6)
    g = K if group == -1 else group
    wd = O.gptq_dequant(packed, scales, zeros)
    want = F.linear(x.cpu().double(), wd.double()).float().to(DEV)
    path = 0 if M * K > 16384 else 2  # long activation blocks do not fit the streaming kernel's shared memory: auto path
    for epi in (0, 2, 3):
        res = f32(M, N, seed=4)
        got = run_linear(lib, x, rows, _lib.LP_W_INT4, N, K, aux0=scales.to(DEV).contiguous(), aux1=zeros.to(DEV).contiguous(),
                         group=g, path=path, packed_aux=True, epi=epi, residual=res)
        torch.testing.assert_close(got, ref_epilogue(want.double(), epi, res.double()).float(), rtol=2e-5, atol=3e-5 * math.sqrt(K) * 0.3)


@pytest.mark.parametrize("M,N,K", [(1, 4096, 4096), (2, 12288, 4096), (1, 4096, 11008), (3, 4544, 18176), (8, 1024, 4096), (16, 512, 2048),
                                   (1, 50688, 4096), (5, 160, 1040)])
def test_linear_stream_bf16_large(lib, M, N, K):
    """Real layer shapes through the streaming family only (path 2): multi-tile persistent CTAs, ragged last stage."""
    w = f32(N, K, seed=1, scale=0.05).bfloat16()
    x, b = f32(M, K, seed=2), f32(N, seed=3)
    rnd = 1 if M > 8 else 0
    if rnd:
        x = bf16r(x)
    ref = F.linear(x.double(), w.double(), b.double())
    path = 2 if (K % 64 == 0 and M * K <= 32768) else 0  # what the streaming kernel covers; else auto (exact CUDA-core kernel)
    for epi in (0, 1, 2, 3):
        res = f32(M, N, seed=4)
        got = run_linear(lib, x, w, _lib.LP_W_BF16, N, K, bias=b, epi=epi, residual=res, path=path, rnd=rnd)
        want = ref_epilogue(ref, epi, res.double()).float()
        if rnd:  # y and epilogue(y) are each rounded to bf16: two half-ulp roundings
            torch.testing.assert_close(got, want, rtol=2 ** -6, atol=2e-2)
        else:
            torch.testing.assert_close(got, want, rtol=2e-5, atol=2e-5 * math.sqrt(K))


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("M,N,K", [(1, 64, 128), (4, 96, 512), (1, 128, 4544 - 64)])
def test_linear_nf4_int8(lib, path, M, N, K):
    w = torch.randn(N, K, generator=torch.Generator().manual_seed(8)) * 0.02
    x = f32(M, K, seed=9)
    packed, absmax = O.nf4_quantize(w)
    want = F.linear(x.cpu().double(), O.nf4_dequantize(packed, absmax, w.shape).double()).float().to(DEV)
    got = run_linear(lib, x, packed.to(DEV), _lib.LP_W_NF4, N, K, aux0=absmax.to(DEV), group=64, path=path)
    torch.testing.assert_close(got, want, rtol=2e-5, atol=1e-5 * math.sqrt(K))
    cb, scb = O.int8_quantize(w)
    want = F.linear(x.cpu().double(), O.int8_dequantize(cb, scb).double()).float().to(DEV)
    got = run_linear(lib, x, cb.to(DEV), _lib.LP_W_INT8, N, K, aux0=(scb / 127.0).to(DEV), path=path)
    torch.testing.assert_close(got, want, rtol=2e-5, atol=1e-5 * math.sqrt(K))


@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("rows,E", [(1, 64), (5, 512), (2, 4544), (3, 8192)])
def test_norm(lib, kind, rows, E):
    x, w, b = f32(rows, E, seed=1), 1 + 0.1 * f32(E, seed=2), 0.1 * f32(E, seed=3)
    y = torch.empty_like(x)
    _lib.check(lib.lp_norm(kind, x.data_ptr(), w.data_ptr(), b.data_ptr() if kind == 0 else None, 1e-5, y.data_ptr(), rows, E, 0,
                           stream()))
    want = F.layer_norm(x, (E,), w, b, 1e-5) if kind == 0 else O.rms_norm(x, w, 1e-5)
    torch.testing.assert_close(y, want, rtol=1e-5, atol=1e-5)
    # bf16-faithful mode against the reference ops evaluated in bf16 on the same (bf16-valued) inputs
    xb, wb, bb = bf16r(x), bf16r(w), bf16r(b)
    _lib.check(lib.lp_norm(kind, xb.data_ptr(), wb.data_ptr(), bb.data_ptr() if kind == 0 else None, 1e-5, y.data_ptr(), rows, E, 1,
                           stream()))
    if kind == 0:
        want = F.layer_norm(xb.bfloat16(), (E,), wb.bfloat16(), bb.bfloat16(), 1e-5).float()
    else:
        want = O.rms_norm(xb.bfloat16(), wb.bfloat16(), 1e-5).float()
    assert (y != want).float().mean() < 0.02  # identical up to rare 1-ulp flips from the reduction order
    torch.testing.assert_close(y, want, rtol=2 ** -7, atol=1e-6)


@pytest.mark.parametrize("wdtype", [torch.float32, torch.bfloat16])
def test_embed(lib, wdtype):
    V, E = 300, 96
    wte = f32(V, E, seed=1).to(wdtype)
    for idt in (torch.int32, torch.int64):
        idx = torch.randint(0, V, (7,), device=DEV).to(idt)
        out = torch.empty(7, E, device=DEV)
        _lib.check(lib.lp_embed(idx.data_ptr(), int(idt == torch.int64), None, wte.data_ptr(), int(wdtype == torch.bfloat16),
                                out.data_ptr(), 7, E, 0, stream()))
        assert torch.equal(out, wte[idx.long()].float())
    off = torch.tensor([3], dtype=torch.int32, device=DEV)
    idx = torch.randint(0, V, (9,), device=DEV, dtype=torch.int32)
    out = torch.empty(1, E, device=DEV)
    _lib.check(lib.lp_embed(idx.data_ptr(), 0, off.data_ptr(), wte.data_ptr(), int(wdtype == torch.bfloat16), out.data_ptr(), 1, E, 0,
                            stream()))
    assert torch.equal(out[0], wte[idx[3].long()].float())


@pytest.mark.parametrize("kvdt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,T,H,G,hs,n_elem", [(1, 1, 8, 8, 64, 16), (2, 5, 8, 2, 128, 128), (1, 3, 71, 1, 64, 64), (3, 4, 4, 4, 16, 4),
                                              (2, 6, 32, 32, 128, 32), (1, 7, 6, 3, 20, 8)])  # 16-byte path and the scalar path (n_elem % 8)
def test_rope_kv_append(lib, kvdt, B, T, H, G, hs, n_elem):
    max_seq, block = 16, 64
    qpk = H // G
    qkv = f32(B * T, (H + 2 * G) * hs, seed=1)
    cos, sin = O.rope_tables(block, n_elem, torch.float32)
    cos, sin = cos.to(DEV).contiguous(), sin.to(DEV).contiguous()
    pos = torch.tensor([7 + t for t in range(T)], dtype=torch.int32, device=DEV)
    q_out = torch.empty(B * T, H * hs, device=DEV)
    kc = torch.zeros(B, G, max_seq, hs, device=DEV, dtype=kvdt)
    vc = torch.zeros_like(kc)
    _lib.check(lib.lp_rope_kv_append(qkv.data_ptr(), cos.data_ptr(), sin.data_ptr(), pos.data_ptr(), q_out.data_ptr(), kc.data_ptr(),
                                     vc.data_ptr(), int(kvdt == torch.bfloat16), B, T, H, G, hs, n_elem, max_seq, 0, stream()))
    # restatement of model.py:208-232 on the same tensor
    v5 = qkv.view(B, T, G, qpk + 2, hs).permute(0, 2, 3, 1, 4)
    q, k, v = v5.split((qpk, 1, 1), dim=2)
    c, s = cos[pos.long()], sin[pos.long()]
    rot = lambda z: torch.cat((O.rotate(z[..., :n_elem], c, s), z[..., n_elem:]), dim=-1)  # noqa: E731
    q, k = rot(q), rot(k)
    want_q = q.permute(0, 3, 1, 2, 4).reshape(B * T, H * hs)
    assert torch.equal(q_out, want_q)  # same fp32 op order: bit-exact
    want_k = torch.zeros(B, G, max_seq, hs, device=DEV)
    want_v = torch.zeros_like(want_k)
    want_k[:, :, pos.long()] = k[:, :, 0]
    want_v[:, :, pos.long()] = v[:, :, 0]
    assert torch.equal(kc.float(), want_k.to(kvdt).float()) and torch.equal(vc.float(), want_v.to(kvdt).float())


@pytest.mark.parametrize("kvdt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,T,H,G,hs,max_seq,p0", [
    (1, 1, 8, 8, 64, 64, 0), (1, 1, 8, 8, 64, 64, 63), (2, 1, 32, 32, 128, 2048, 1500), (1, 1, 71, 1, 64, 512, 300),
    (2, 1, 64, 8, 128, 1024, 1023), (1, 7, 4, 2, 16, 32, 0), (3, 5, 8, 8, 32, 40, 9), (1, 1, 8, 8, 64, 48, 200),
    (1, 3, 2, 2, 4, 25, 3), (1, 2, 4, 4, 2, 25, 6), (2, 1, 3, 1, 6, 20, 11)])
def test_attention_against_sdpa(lib, kvdt, B, T, H, G, hs, max_seq, p0):
    """lp_attn_decode against masked SDPA over the zero-filled cache, as the reference runs it (model.py:91-92, 273-275).
    p0 >= max_seq exercises the ring (sliding-window) case: all max_seq slots are attended."""
    qpk = H // G
    q = f32(B * T, H * hs, seed=1)
    pos = torch.arange(p0, p0 + T, dtype=torch.int32, device=DEV)
    kc = torch.zeros(B, G, max_seq, hs, device=DEV, dtype=kvdt)
    vc = torch.zeros_like(kc)
    n_valid = min(p0 + T, max_seq)
    kc[:, :, :n_valid] = f32(B, G, n_valid, hs, seed=2).to(kvdt)
    vc[:, :, :n_valid] = f32(B, G, n_valid, hs, seed=3).to(kvdt)
    out = torch.empty(B * T, H * hs, device=DEV)
    ws = torch.empty(lib.lp_attn_workspace_bytes(B, T, H, hs, max_seq) + 16, dtype=torch.uint8, device=DEV)
    scale = 1.0 / math.sqrt(hs)
    _lib.check(lib.lp_attn_decode(q.data_ptr(), kc.data_ptr(), vc.data_ptr(), int(kvdt == torch.bfloat16), pos.data_ptr(), out.data_ptr(),
                                  ws.data_ptr(), ws.numel(), B, T, H, G, hs, max_seq, scale, 0, stream()))
    qq = q.view(B, T, H, hs).transpose(1, 2).double()
    kk = kc.double().repeat_interleave(qpk, dim=1)
    vv = vc.double().repeat_interleave(qpk, dim=1)
    kpos = torch.arange(max_seq, device=DEV)
    mask = kpos[None, :] <= torch.clamp(pos.long(), max=max_seq - 1)[:, None]  # row t sees min(pos+1, max_seq) slots
    want = F.scaled_dot_product_attention(qq, kk, vv, attn_mask=mask[None, None], scale=scale)
    want = want.transpose(1, 2).reshape(B * T, H * hs).float()
    torch.testing.assert_close(out, want, rtol=1e-4, atol=2e-5)


@pytest.mark.parametrize("B,H,G,hs,n_elem,max_seq,p0", [
    (1, 8, 8, 64, 16, 64, 0), (1, 8, 8, 64, 16, 64, 63), (2, 32, 32, 128, 32, 2048, 1500), (1, 71, 1, 64, 64, 512, 300),
    (2, 64, 8, 128, 128, 1024, 1023), (1, 32, 32, 128, 128, 2048, 2047), (1, 8, 8, 64, 64, 48, 200), (3, 128, 8, 64, 64, 200, 130),
    (1, 64, 2, 128, 128, 300, 77), (32, 32, 32, 128, 32, 256, 100),
    # more (batch, group) CTAs than the chip holds at once (no split; a wave-aware split count was measured SLOWER at B = 32:
    # 5.65 -> 5.78 ms per step — a partially filled tail wave still saturates HBM, the merge pass does not pay)
    (19, 16, 16, 128, 128, 1024, 1000), (32, 32, 32, 128, 32, 2048, 2047), (40, 16, 16, 64, 16, 512, 300)])
def test_attention_decode_fused(lib, B, H, G, hs, n_elem, max_seq, p0):
    """lp_attn_decode_fused (RoPE + append + split-K tensor-core attention + merge in one launch) against the reference
    op sequence (model.py:208-249) in float64 on the same bf16 cache.  p0 >= max_seq: ring / sliding-window case."""
    qpk = H // G
    qkv = f32(B, (H + 2 * G) * hs, seed=1)
    cos, sin = O.rope_tables(max(p0 + 1, 64), n_elem, torch.float32)
    cos, sin = cos.to(DEV).contiguous(), sin.to(DEV).contiguous()
    pos = torch.tensor([p0], dtype=torch.int32, device=DEV)
    kc = torch.zeros(B, G, max_seq, hs, device=DEV, dtype=torch.bfloat16)
    vc = torch.zeros_like(kc)
    n_old = min(p0, max_seq)
    kc[:, :, :n_old] = f32(B, G, n_old, hs, seed=2).bfloat16()
    vc[:, :, :n_old] = f32(B, G, n_old, hs, seed=3).bfloat16()
    k_before, v_before = kc.clone(), vc.clone()
    out = torch.full((B, H * hs), float("nan"), device=DEV)
    ws = torch.zeros(lib.lp_attn_fused_workspace_bytes(B, H, G, hs, max_seq) + 16, dtype=torch.uint8, device=DEV)
    scale = 1.0 / math.sqrt(hs)
    for rep in range(2):  # twice: the ticket area must be left zeroed
        kc.copy_(k_before)
        vc.copy_(v_before)
        _lib.check(lib.lp_attn_decode_fused(qkv.data_ptr(), cos.data_ptr(), sin.data_ptr(), pos.data_ptr(), out.data_ptr(),
                                            kc.data_ptr(), vc.data_ptr(), _lib.LP_BF16, ws.data_ptr(), ws.numel(), B, H, G, hs, n_elem,
                                            max_seq, scale, 0, stream()), "lp_attn_decode_fused")
    torch.cuda.synchronize()
    v5 = qkv.view(B, 1, G, qpk + 2, hs).permute(0, 2, 3, 1, 4)
    q, k, v = v5.split((qpk, 1, 1), dim=2)
    c, s = cos[pos.long()], sin[pos.long()]
    rot = lambda z: torch.cat((O.rotate(z[..., :n_elem], c, s), z[..., n_elem:]), dim=-1)  # noqa: E731
    q, k = rot(q), rot(k)
    slot = p0 % max_seq
    want_k, want_v = k_before.clone(), v_before.clone()
    want_k[:, :, slot] = k[:, :, 0, 0].bfloat16()
    want_v[:, :, slot] = v[:, :, 0, 0].bfloat16()
    assert torch.equal(kc, want_k) and torch.equal(vc, want_v)
    n_valid = min(p0 + 1, max_seq)
    qq = q.reshape(B, H, 1, hs).double()
    kk = want_k[:, :, :n_valid].double().repeat_interleave(qpk, dim=1)
    vv = want_v[:, :, :n_valid].double().repeat_interleave(qpk, dim=1)
    want = F.scaled_dot_product_attention(qq, kk, vv, scale=scale).reshape(B, H * hs).float()
    torch.testing.assert_close(out, want, rtol=1e-4, atol=2e-5)
    assert int(ws[:4 * B * G].view(torch.int32).abs().sum()) == 0  # tickets reset for the next launch


def test_sample_greedy_and_topk(lib):
    V = 50304
    logits = f32(2, V, seed=1)
    logits[0, 777] = logits[0, 12345] = 9.0  # exact tie -> lowest index
    tok = torch.full((2,), -1, dtype=torch.int32, device=DEV)
    step = torch.zeros(1, dtype=torch.int32, device=DEV)
    _lib.check(lib.lp_sample(logits.data_ptr(), 2, V, 1.0, 1, 1234, step.data_ptr(), tok.data_ptr(), None, None, stream()))
    assert tok.tolist() == [777, int(logits[1].argmax())]
    # top-k: every draw must come from the k best logits (ties with the k-th value survive, generate/base.py:139-141)
    k = 5
    lg = f32(1, V, seed=2)
    kth = torch.topk(lg[0], k).values[-1]
    lg[0, 4242] = kth  # a tie with the k-th value must stay eligible
    allowed = set((lg[0] >= kth).nonzero().flatten().tolist())
    assert len(allowed) == k + 1
    seen = set()
    one = torch.zeros(1, dtype=torch.int32, device=DEV)
    for i in range(300):
        _lib.check(lib.lp_sample(lg.data_ptr(), 1, V, 2.0, k, 99, step.data_ptr(), one.data_ptr(), None, None, stream()))
        seen.add(int(one))
    assert seen <= allowed and len(seen) == k + 1 and int(step) == 301  # the 2-row greedy launch above advanced it too
    # device-side append used by the captured decode step
    seq = torch.zeros(8, dtype=torch.int32, device=DEV)
    pos = torch.tensor([2], dtype=torch.int32, device=DEV)
    _lib.check(lib.lp_sample(lg.data_ptr(), 1, V, 1.0, 1, 0, step.data_ptr(), one.data_ptr(), seq.data_ptr(), pos.data_ptr(), stream()))
    assert int(pos) == 3 and int(seq[3]) == int(lg[0].argmax()) == int(one)


def test_sample_bf16_logits_match_fp32_of_the_same_values(lib):
    """lp_sample_bf16 (the logits GPT.forward returns for a bf16 checkpoint, no conversion pass): same tokens as lp_sample on the
    same values widened to fp32 — greedy incl. the lowest-index tie break, and the seeded top-k draws one for one."""
    import lit_parrot_b200 as lp

    V = 50304
    lg16 = f32(3, V, seed=5).bfloat16()
    lg16[0, 100] = lg16[0, 40000] = 12.0  # exact tie -> lowest index
    lg32 = lg16.float()
    a = torch.zeros(3, dtype=torch.int32, device=DEV)
    b = torch.zeros_like(a)
    for k, temp in ((1, 1.0), (7, 0.8), (0, 1.3)):
        sa, sb = torch.zeros(1, dtype=torch.int32, device=DEV), torch.zeros(1, dtype=torch.int32, device=DEV)
        for _ in range(4):
            _lib.check(lib.lp_sample_bf16(lg16.data_ptr(), 3, V, temp, k, 77, sa.data_ptr(), a.data_ptr(), None, None, stream()))
            _lib.check(lib.lp_sample(lg32.data_ptr(), 3, V, temp, k, 77, sb.data_ptr(), b.data_ptr(), None, None, stream()))
            assert torch.equal(a, b) and int(sa) == int(sb)
        if k == 1:
            assert int(a[0]) == 100
    # the Python entry point takes either dtype
    assert torch.equal(lp.sample(lg16, 1.0, 1), lp.sample(lg32, 1.0, 1))


def test_sample_batched_step_advances(lib):
    """Multi-row launches advance the Philox step once per launch (the noise is keyed by (index, step, row)): replaying the same
    launch on the same logits must draw fresh noise every time, and a rewound step must reproduce the earlier draw."""
    V, rows = 4096, 8
    lg = torch.zeros(rows, V, device=DEV)  # uniform: the draw is the arg-max of the noise alone
    tok = torch.zeros(rows, dtype=torch.int32, device=DEV)
    step = torch.tensor([5], dtype=torch.int32, device=DEV)
    draws = []
    for i in range(6):
        _lib.check(lib.lp_sample(lg.data_ptr(), rows, V, 1.0, 0, 42, step.data_ptr(), tok.data_ptr(), None, None, stream()))
        draws.append(tok.cpu().clone())
        assert int(step) == 6 + i
    assert all(not torch.equal(draws[i], draws[i + 1]) for i in range(5))
    assert len({tuple(d.tolist()) for d in draws}) == 6
    step.fill_(7)  # the third launch again
    _lib.check(lib.lp_sample(lg.data_ptr(), rows, V, 1.0, 0, 42, step.data_ptr(), tok.data_ptr(), None, None, stream()))
    assert torch.equal(tok.cpu(), draws[2])
    # greedy rows advance the step as well (it keys nothing there, but the counter stays in step with the launches)
    _lib.check(lib.lp_sample(lg.data_ptr(), rows, V, 1.0, 1, 42, step.data_ptr(), tok.data_ptr(), None, None, stream()))
    assert int(step) == 9


def test_sample_distribution(lib):
    """temperature sampling follows softmax(logits / T) (chi-square over 20k draws on a small vocabulary)."""
    V, n, T = 12, 20000, 0.7
    lg = f32(1, V, seed=4)
    p = F.softmax(lg[0].double() / T, dim=-1).cpu().numpy()
    rows = 500
    big = lg.repeat(rows, 1).contiguous()
    tok = torch.zeros(rows, dtype=torch.int32, device=DEV)
    step = torch.zeros(1, dtype=torch.int32, device=DEV)
    counts = np.zeros(V)
    for i in range(n // rows):
        step.fill_(i)
        _lib.check(lib.lp_sample(big.data_ptr(), rows, V, T, 0, 7, step.data_ptr(), tok.data_ptr(), None, None, stream()))
        counts += np.bincount(tok.cpu().numpy(), minlength=V)
    chi2 = float(((counts - n * p) ** 2 / (n * p)).sum())
    assert chi2 < 45.0, chi2  # 11 dof: P(chi2 > 45) ~ 5e-6


@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("fmt", ["bf16", "int4"])
@pytest.mark.parametrize("M,N,K", [(1, 64, 128), (2, 256, 512), (4, 96, 1024), (1, 4672, 4544)])
def test_norm_linear_fused(lib, kind, fmt, M, N, K):
    """lp_norm_linear (norm as the GEMV prologue) == lp_norm followed by lp_linear == torch restatement."""
    x, nw, nb = f32(M, K, seed=1), 1 + 0.1 * f32(K, seed=2), 0.1 * f32(K, seed=3)
    xn = (F.layer_norm(x, (K,), nw, nb, 1e-5) if kind == 0 else O.rms_norm(x, nw, 1e-5)).double()
    p = lambda a: None if a is None else a.data_ptr()  # noqa: E731
    if fmt == "bf16":
        w = f32(N, K, seed=4, scale=0.05).bfloat16()
        rec = LpWeight(w.data_ptr(), None, None, None, None, _lib.LP_W_BF16, N, K, 0, 0, 0)
        want = F.linear(xn, w.double()).float()
    else:
        wf = torch.randn(N, K, generator=torch.Generator().manual_seed(5)) * 0.02
        packed, scales, zeros = O.gptq_rtn_quantize(wf, 128)
        src = torch.empty((K // 2, N), dtype=torch.uint8, device=DEV).t()
        src.copy_(packed)
        rows = torch.empty((N, lib.lp_int4_row_bytes(K)), dtype=torch.uint8, device=DEV)
        _lib.check(lib.lp_repack_gptq_int4(src.data_ptr(), rows.data_ptr(), N, K, stream()))
        sc, ze = scales.to(DEV).contiguous(), zeros.to(DEV).contiguous()
        from lit_parrot_b200.quantize import tile_major_aux
        aux2, flags = tile_major_aux(sc, ze)
        rec = LpWeight(rows.data_ptr(), sc.data_ptr(), ze.data_ptr(), aux2.data_ptr(), None, _lib.LP_W_INT4, N, K, 128, flags, 0)
        want = F.linear(xn, O.gptq_dequant(packed, scales, zeros, tile_cols=128).double().to(DEV)).float()
    out = torch.full((M, N), float("nan"), device=DEV)
    rc = lib.lp_norm_linear(kind, nw.data_ptr(), p(nb if kind == 0 else None), 1e-5, x.data_ptr(), M, ctypes.byref(rec), 0, None,
                            out.data_ptr(), 0, stream())
    _lib.check(rc, "lp_norm_linear")
    torch.cuda.synchronize()
    torch.testing.assert_close(out, want, rtol=3e-5, atol=3e-5 * math.sqrt(K))
    # the fusion must decline (not mis-compute) what it does not cover
    assert lib.lp_norm_linear(kind, nw.data_ptr(), None, 1e-5, x.data_ptr(), M, ctypes.byref(rec), 0, None, out.data_ptr(), 1, stream()) == -2


@pytest.mark.parametrize("nterms", [1, 2, 3])
@pytest.mark.parametrize("M,N,K", [(16, 128, 64), (128, 256, 512), (200, 384, 4096), (333, 4608, 4544), (2048, 512, 1024), (9, 1280, 8192),
                                   (150, 4544, 1024), (70, 200, 136), (32, 4096, 4096), (33, 4672, 4544), (64, 200, 136), (48, 16384, 512),
                                   (1024, 4864, 256), (777, 6400, 320), (512, 9736, 128)])  # the last three: CTA-pair (cta_group::2) kernel
def test_gemm_bf16_tc(lib, nterms, M, N, K):
    """lp_split_bf16 + lp_gemm_bf16_tc (TMA + tcgen05.mma, accumulator in tensor memory) against float64 F.linear.
    nterms = 1: bf16 activations (the reference's bf16-true matmul inputs); 2 / 3: fp32-activation accuracy."""
    lib.lp_set_gemm_pair(1 if (M, N, K) in ((1024, 4864, 256), (777, 6400, 320), (512, 9736, 128)) else 0)
    x = f32(M, K, seed=1)
    w = f32(N, K, seed=2, scale=0.05).bfloat16()
    bias = f32(N, seed=3)
    res = f32(M, N, seed=4)
    terms = torch.empty(nterms, M, K, dtype=torch.bfloat16, device=DEV)
    _lib.check(lib.lp_split_bf16(x.data_ptr(), terms.data_ptr(), M, K, nterms, -1, None, None, 0.0, 0, stream()), "lp_split_bf16")
    xe = terms.double().sum(0)  # what the tensor cores see
    if nterms == 3:
        assert torch.equal(xe.float(), x)  # three bf16 terms carry all 24 bits
    ref = F.linear(xe, w.double(), bias.double())
    tol = dict(rtol=2e-5, atol=2e-5 * math.sqrt(K))
    for epi in (0, 1, 2, 3):
        nout = N // 2 if epi == 2 else N
        out = torch.full((M, nout), float("nan"), device=DEV)
        out_t = torch.zeros(2, M, nout, dtype=torch.bfloat16, device=DEV)
        _lib.check(lib.lp_gemm_bf16_tc(terms.data_ptr(), nterms, M, w.data_ptr(), N, K, bias.data_ptr(), epi, res.data_ptr(),
                                       out.data_ptr(), out_t.data_ptr(), 2, 0, stream()), "lp_gemm_bf16_tc")
        torch.cuda.synchronize()
        want = ref_epilogue(ref, epi, res.double()).float()
        if epi == 2:  # silu(a) * b multiplies the absolute error of a by |b| (up to ~10 here)
            torch.testing.assert_close(out, want, rtol=5e-4, atol=2e-4 * math.sqrt(K))
        else:
            torch.testing.assert_close(out, want, **tol)
        # the bf16 split of the result (operand of the next GEMM): hi + lo reproduces it to 16 bits
        torch.testing.assert_close(out_t.float().sum(0), out, rtol=2 ** -15, atol=1e-30)
    lib.lp_set_gemm_pair(1)  # the library default
    # fused norm in the splitter
    nw, nb = 1 + 0.1 * f32(K, seed=5), 0.1 * f32(K, seed=6)
    for kind in (0, 1):
        _lib.check(lib.lp_split_bf16(x.data_ptr(), terms.data_ptr(), M, K, nterms, kind, nw.data_ptr(), nb.data_ptr() if kind == 0 else None,
                                     1e-5, 0, stream()), "lp_split_bf16")
        want = F.layer_norm(x, (K,), nw, nb, 1e-5) if kind == 0 else O.rms_norm(x, nw, 1e-5)
        torch.testing.assert_close(terms.float().sum(0), want, rtol=2 ** (-8 * nterms + 1), atol=1e-5)


@pytest.mark.parametrize("M,N,K", [(32, 4096, 4096), (9, 512, 16384), (64, 4544, 18176), (17, 256, 192), (32, 12288, 4096),
                                   (32, 4096, 16384), (24, 6144, 2048), (48, 8192, 8192)])
def test_gemm_bf16_tc_inplace_residual_split_k(lib, M, N, K):
    """Decode batches (M <= 64): swap-AB kernel; x += u . W^T + b in place is split along K — stream-K: the (tile, K-stage)
    sequence is cut into one equal range per SM, a range may straddle two tiles — over CTAs that accumulate atomically (the
    summation order of the partial sums is not fixed: tolerance, not bit-equality)."""
    u = f32(M, K, seed=11)
    w = f32(N, K, seed=12, scale=0.05).bfloat16()
    bias = f32(N, seed=13)
    x0 = f32(M, N, seed=14)
    terms = torch.empty(2, M, K, dtype=torch.bfloat16, device=DEV)
    _lib.check(lib.lp_split_bf16(u.data_ptr(), terms.data_ptr(), M, K, 2, -1, None, None, 0.0, 0, stream()), "lp_split_bf16")
    want = (x0.double() + F.linear(terms.double().sum(0), w.double(), bias.double())).float()
    x = x0.clone()
    _lib.check(lib.lp_gemm_bf16_tc(terms.data_ptr(), 2, M, w.data_ptr(), N, K, bias.data_ptr(), 3, x.data_ptr(), x.data_ptr(), None, 0, 0,
                                   stream()), "lp_gemm_bf16_tc")
    torch.cuda.synchronize()
    torch.testing.assert_close(x, want, rtol=2e-5, atol=2e-5 * math.sqrt(K))


@pytest.mark.parametrize("fmt", ["int4", "nf4", "int8"])
def test_dequant_bf16(lib, fmt):
    N, K = 96, 512
    w = torch.randn(N, K, generator=torch.Generator().manual_seed(8)) * 0.02
    out = torch.empty(N, K, dtype=torch.bfloat16, device=DEV)
    p = lambda a: a.data_ptr()  # noqa: E731
    if fmt == "int4":
        packed, scales, zeros = O.gptq_rtn_quantize(w, 128)
        src = torch.empty((K // 2, N), dtype=torch.uint8, device=DEV).t()
        src.copy_(packed)
        rows = torch.empty((N, lib.lp_int4_row_bytes(K)), dtype=torch.uint8, device=DEV)
        _lib.check(lib.lp_repack_gptq_int4(src.data_ptr(), rows.data_ptr(), N, K, stream()))
        sc, ze = scales.to(DEV).contiguous(), zeros.to(DEV).contiguous()
        rec = LpWeight(p(rows), p(sc), p(ze), None, None, _lib.LP_W_INT4, N, K, 128, 0, 0)
        want = O.gptq_dequant(packed, scales, zeros).bfloat16()
    elif fmt == "nf4":
        packed, absmax = O.nf4_quantize(w)
        pk, am = packed.to(DEV), absmax.to(DEV)
        rec = LpWeight(p(pk), p(am), None, None, None, _lib.LP_W_NF4, N, K, 64, 0, 0)
        want = O.nf4_dequantize(packed, absmax, w.shape).bfloat16()
    else:
        cb, scb = O.int8_quantize(w)
        cbd, sc = cb.to(DEV), (scb / 127.0).to(DEV)
        rec = LpWeight(p(cbd), p(sc), None, None, None, _lib.LP_W_INT8, N, K, 0, 0, 0)
        want = O.int8_dequantize(cb, scb).bfloat16()
    _lib.check(lib.lp_dequant_bf16(ctypes.byref(rec), out.data_ptr(), stream()), "lp_dequant_bf16")
    torch.cuda.synchronize()
    torch.testing.assert_close(out.float().cpu(), want.float(), rtol=2 ** -7, atol=1e-8)


@pytest.mark.parametrize("path", [1, 2], ids=["mma_sync", "tcgen05"])
@pytest.mark.parametrize("rnd", [0, 1])
@pytest.mark.parametrize("B,T,H,G,hs,max_seq,p0", [(1, 16, 8, 8, 64, 64, 0), (2, 100, 32, 32, 128, 256, 0), (1, 300, 71, 1, 64, 512, 0),
                                                   (1, 130, 64, 8, 128, 1024, 500), (2, 65, 8, 2, 64, 200, 7), (1, 2048, 4, 4, 128, 2048, 0),
                                                   (1, 128, 2, 2, 128, 128, 0), (1, 129, 3, 1, 64, 640, 511)])
def test_attention_prefill(lib, rnd, B, T, H, G, hs, max_seq, p0, path):
    """lp_attn_prefill over the bf16 cache — both kernels: FlashAttention-2 style mma.sync, and tcgen05.mma with the S / O
    accumulators in tensor memory (csrc/attention_tc.cu) — against masked SDPA in float64."""
    lib.lp_set_attn_prefill_path(path)
    qpk = H // G
    q = f32(B * T, H * hs, seed=1)
    if rnd:
        q = bf16r(q)
    pos = torch.arange(p0, p0 + T, dtype=torch.int32, device=DEV)
    kc = torch.zeros(B, G, max_seq, hs, device=DEV, dtype=torch.bfloat16)
    vc = torch.zeros_like(kc)
    n_valid = p0 + T
    kc[:, :, :n_valid] = f32(B, G, n_valid, hs, seed=2).bfloat16()
    vc[:, :, :n_valid] = f32(B, G, n_valid, hs, seed=3).bfloat16()
    out = torch.full((B * T, H * hs), float("nan"), device=DEV)
    scale = 1.0 / math.sqrt(hs)
    try:
        _lib.check(lib.lp_attn_prefill(q.data_ptr(), kc.data_ptr(), vc.data_ptr(), _lib.LP_BF16, pos.data_ptr(), out.data_ptr(), B, T, H, G, hs,
                                       max_seq, scale, rnd, stream()), "lp_attn_prefill")
        torch.cuda.synchronize()
    finally:
        lib.lp_set_attn_prefill_path(0)
    qq = q.view(B, T, H, hs).transpose(1, 2).double()
    kk = kc[:, :, :n_valid].double().repeat_interleave(qpk, dim=1)
    vv = vc[:, :, :n_valid].double().repeat_interleave(qpk, dim=1)
    mask = torch.arange(n_valid, device=DEV)[None, :] <= pos.long()[:, None]
    want = F.scaled_dot_product_attention(qq, kk, vv, attn_mask=mask[None, None], scale=scale).transpose(1, 2).reshape(B * T, H * hs).float()
    if rnd:  # P and the output are rounded to bf16, as in the reference's bf16 SDPA
        torch.testing.assert_close(out, want, rtol=2 ** -6, atol=2e-2)
    else:
        torch.testing.assert_close(out, want, rtol=1e-4, atol=2e-5)
