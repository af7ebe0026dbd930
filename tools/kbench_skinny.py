"""Decode-batch projections (M = 9..64) on lp_gemm_bf16_tc (swap-AB kernel): GB/s of weight streaming.  Tuning aid."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lit_parrot_b200 import _lib  # noqa: E402

DEV = torch.device("cuda", 0)
lib = _lib.init(0)
st = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731


def timeit(fn, iters=40, warm=5):
    for i in range(warm):
        fn(i)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


print(f"{'M,N,K':24s} {'terms':>5s} {'epi':>4s} {'us':>8s} {'GB/s':>7s}  torch.matmul us")
for M in [int(v) for v in os.environ.get("KBENCH_M", "16,32,64").split(",")]:
    shapes = [(12288, 4096, 0), (16384, 4096, 1), (4096, 4096, 3), (4096, 16384, 3), (50688, 4096, 0)]
    if os.environ.get("KBENCH_SHAPES"):  # "N,K,epi;N,K,epi;..."
        shapes = [tuple(int(v) for v in t.split(",")) for t in os.environ["KBENCH_SHAPES"].split(";")]
    for N, K, epi in shapes:
        nt = int(os.environ.get("KBENCH_TERMS", "2"))
        x = torch.randn(M, K, device=DEV)
        w = [torch.randn(N, K, device=DEV, dtype=torch.bfloat16) * 0.02 for _ in range(6)]  # 6 x >= 33 MB: beyond L2 in rotation
        out = torch.zeros(M, N, device=DEV)
        terms = torch.empty(nt, M, K, dtype=torch.bfloat16, device=DEV)
        _lib.check(lib.lp_split_bf16(x.data_ptr(), terms.data_ptr(), M, K, nt, -1, None, None, 0.0, 0, st()))
        res = out.data_ptr() if epi == 3 else None
        fn = lambda i: lib.lp_gemm_bf16_tc(terms.data_ptr(), nt, M, w[i % 6].data_ptr(), N, K, None, epi, res, out.data_ptr(), None, 0, 0, st())  # noqa: E731
        assert fn(0) == 0
        us = timeit(fn)
        xb = x.bfloat16()
        ust = timeit(lambda i: torch.matmul(xb, w[i % 6].t()))
        print(f"{str((M, N, K)):24s} {nt:5d} {epi:4d} {us:8.1f} {N * K * 2 / us / 1e3:7.0f}  {ust:8.1f}")
