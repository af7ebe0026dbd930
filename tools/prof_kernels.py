"""A handful of launches of the hot kernels at real layer shapes, for `ncu` (see profiles/README.md).  Not a benchmark."""
import ctypes
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lit_parrot_b200 import _lib  # noqa: E402
from lit_parrot_b200._lib import LpWeight  # noqa: E402
from lit_parrot_b200.quantize import tile_major_aux  # noqa: E402

DEV = torch.device("cuda", 0)
lib = _lib.init(0)
st = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731
what = sys.argv[1:] or ["bf16", "int4", "attn"]
reps = 3

if "bf16" in what or "int4" in what:
    N, K = 16384, 4096
    x = torch.randn(1, K, device=DEV)
    out = torch.empty(1, N, device=DEV)
    lib.lp_set_linear_path(2)
    if "bf16" in what:
        ws = [torch.randn(N, K, device=DEV, dtype=torch.bfloat16) * 0.02 for _ in range(3)]
        for i in range(reps):
            rec = LpWeight(ws[i % 3].data_ptr(), None, None, None, None, _lib.LP_W_BF16, N, K, 0, 0, 0)
            _lib.check(lib.lp_linear(x.data_ptr(), 1, ctypes.byref(rec), 0, None, out.data_ptr(), 0, st()))
    if "int4" in what:
        N = 22016
        out = torch.empty(1, N, device=DEV)
        rb = lib.lp_int4_row_bytes(K)
        keep = []
        for i in range(reps):
            w = torch.randint(0, 256, (N, rb), device=DEV, dtype=torch.uint8)
            sc = (torch.rand(N, K // 128, device=DEV) * 0.01).bfloat16().float()
            ze = torch.full((N, K // 128), 8.0, device=DEV)
            aux2, flags = tile_major_aux(sc, ze)
            keep += [w, sc, ze, aux2]
            rec = LpWeight(w.data_ptr(), sc.data_ptr(), ze.data_ptr(), aux2.data_ptr(), None, _lib.LP_W_INT4, N, K, 128, flags, 0)
            _lib.check(lib.lp_linear(x.data_ptr(), 1, ctypes.byref(rec), 0, None, out.data_ptr(), 0, st()))
    lib.lp_set_linear_path(0)
if "fma" in what:
    N, K = 16384, 4096
    x = torch.randn(1, K, device=DEV)
    out = torch.empty(1, N, device=DEV)
    lib.lp_set_linear_path(1)
    ws = [torch.randn(N, K, device=DEV, dtype=torch.bfloat16) * 0.02 for _ in range(3)]
    for i in range(reps):
        rec = LpWeight(ws[i % 3].data_ptr(), None, None, None, None, _lib.LP_W_BF16, N, K, 0, 0, 0)
        _lib.check(lib.lp_linear(x.data_ptr(), 1, ctypes.byref(rec), 0, None, out.data_ptr(), 0, st()))
    lib.lp_set_linear_path(0)
if "attn" in what:
    for B, H, G, hs, ctx in [(1, 32, 32, 128, 2048), (1, 71, 1, 64, 2048)]:
        kc = [torch.randn(B, G, ctx, hs, device=DEV).bfloat16() for _ in range(reps)]
        vc = [torch.randn(B, G, ctx, hs, device=DEV).bfloat16() for _ in range(reps)]
        qkv = torch.randn(B, (H + 2 * G) * hs, device=DEV)
        cos = torch.randn(ctx, hs, device=DEV)
        out = torch.empty(B, H * hs, device=DEV)
        pos = torch.tensor([ctx - 1], dtype=torch.int32, device=DEV)
        wsf = torch.zeros(lib.lp_attn_fused_workspace_bytes(B, H, G, hs, ctx) + 16, dtype=torch.uint8, device=DEV)
        for i in range(reps):
            _lib.check(lib.lp_attn_decode_fused(qkv.data_ptr(), cos.data_ptr(), cos.data_ptr(), pos.data_ptr(), out.data_ptr(),
                                                kc[i].data_ptr(), vc[i].data_ptr(), 1, wsf.data_ptr(), wsf.numel(), B, H, G, hs, hs, ctx,
                                                1 / math.sqrt(hs), 0, st()))
torch.cuda.synchronize()
print("ok")
