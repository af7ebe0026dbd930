"""Prefill attention alone: mma.sync (FlashAttention-2 style) kernel vs the tcgen05 / TMEM kernel.  TFLOP/s counts the causal half."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lit_parrot_b200 import _lib  # noqa: E402

DEV = torch.device("cuda", 0)
lib = _lib.init(0)
st = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731
print(f"{'B,T,H,G,hs':26s} {'mode':>5s} {'kernel':>9s} {'us':>9s} {'TFLOP/s':>8s}")
for B, T, H, G, hs in [(1, 1792, 71, 1, 64), (1, 2048, 32, 32, 128), (1, 2048, 64, 8, 128), (4, 512, 32, 32, 128)]:
    q = torch.randn(B * T, H * hs, device=DEV)
    kc = torch.randn(B, G, T, hs, device=DEV).bfloat16()
    vc = torch.randn(B, G, T, hs, device=DEV).bfloat16()
    out = torch.empty(B * T, H * hs, device=DEV)
    pos = torch.arange(T, dtype=torch.int32, device=DEV)
    flops = 4.0 * B * H * hs * T * (T + 1) / 2
    for rnd in (0, 1):
        for path, name in ((1, "mma.sync"), (2, "tcgen05")):
            lib.lp_set_attn_prefill_path(path)
            fn = lambda: lib.lp_attn_prefill(q.data_ptr(), kc.data_ptr(), vc.data_ptr(), 1, pos.data_ptr(), out.data_ptr(), B, T, H, G, hs, T,  # noqa: E731
                                             1 / math.sqrt(hs), rnd, st())
            assert fn() == 0
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(10):
                fn()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 100
            print(f"{str((B, T, H, G, hs)):26s} {'bf16' if rnd else 'fp32':>5s} {name:>9s} {us:9.1f} {flops / us / 1e6:8.1f}")
lib.lp_set_attn_prefill_path(0)
