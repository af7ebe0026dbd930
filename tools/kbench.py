"""Kernel micro-benchmarks on a B200 (CUDA events, rotating operands larger than L2).  Not a bench.py value: a tuning aid.

    python tools/kbench.py [linear] [attn]
"""
import ctypes
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lit_parrot_b200 import _lib  # noqa: E402
from lit_parrot_b200._lib import LpWeight  # noqa: E402

DEV = torch.device("cuda", 0)
lib = _lib.init(0)
PEAK = 6550.0


def stream():
    return torch.cuda.current_stream().cuda_stream


def timeit(fn, iters=50, warm=5):
    for i in range(warm):
        fn(i)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters  # us


def bench_linear():
    shapes = [("qkv7b", 12288, 4096), ("proj7b", 4096, 4096), ("fc7b(swiglu)", 22016, 4096), ("mlpproj7b", 4096, 11008),
              ("lmhead7b", 32000, 4096), ("fc3b", 16384, 4096), ("falcon qkv", 4672, 4544), ("fc70b/tp8", 7168, 8192)]
    print(f"{'shape':16s} {'fmt':5s} {'M':>2s} {'path':4s} {'us':>8s} {'GB/s':>8s} {'%peak':>6s}")
    for name, N, K in shapes:
        for fmt in ("bf16", "int4"):
            nbuf = max(2, int(400e6 // (N * K * (2 if fmt == "bf16" else 0.5))) + 1)
            nbuf = min(nbuf, 24)
            recs, keep = [], []
            for i in range(nbuf):
                if fmt == "bf16":
                    w = torch.randn(N, K, device=DEV, dtype=torch.bfloat16) * 0.02
                    keep.append(w)
                    recs.append(LpWeight(w.data_ptr(), None, None, None, None, _lib.LP_W_BF16, N, K, 0, 0, 0))
                    nbytes = N * K * 2
                else:
                    rb = lib.lp_int4_row_bytes(K)
                    w = torch.randint(0, 256, (N, rb), device=DEV, dtype=torch.uint8)
                    ng = (K + 127) // 128
                    sc = torch.rand(N, ng, device=DEV) * 0.01
                    ze = torch.full((N, ng), 8.0, device=DEV)
                    keep += [w, sc, ze]
                    from lit_parrot_b200.quantize import tile_major_aux
                    sc = sc.bfloat16().float()
                    aux2, flags = tile_major_aux(sc, ze)
                    keep.append(aux2)
                    recs.append(LpWeight(w.data_ptr(), sc.data_ptr(), ze.data_ptr(), aux2.data_ptr(), None, _lib.LP_W_INT4, N, K, 128, flags, 0))
                    nbytes = N * rb + N * ng * 4
            for M in (1, 2, 4, 8):
                x = torch.randn(M, K, device=DEV)
                out = torch.empty(M, N, device=DEV)
                for path, pname in ((1, "fma"), (2, "strm")):
                    lib.lp_set_linear_path(path)
                    rc = lib.lp_linear(x.data_ptr(), M, ctypes.byref(recs[0]), 0, None, out.data_ptr(), 0, stream())
                    if rc != 0:
                        print(f"{name:16s} {fmt:5s} {M:2d} {pname:4s} unsupported ({rc})")
                        continue
                    us = timeit(lambda i: lib.lp_linear(x.data_ptr(), M, ctypes.byref(recs[i % nbuf]), 0, None, out.data_ptr(), 0, stream()))
                    print(f"{name:16s} {fmt:5s} {M:2d} {pname:4s} {us:8.2f} {nbytes / us / 1e3:8.0f} {100 * nbytes / us / 1e3 / PEAK:6.1f}")
            lib.lp_set_linear_path(0)
            del recs, keep
            torch.cuda.empty_cache()
    # plain streaming read reference: torch sum over a large bf16 buffer
    big = torch.empty(1 << 30, device=DEV, dtype=torch.bfloat16).normal_()
    us = timeit(lambda i: big.sum(), iters=10, warm=2)
    print(f"torch.sum over 2 GiB: {us:.0f} us = {big.numel() * 2 / us / 1e3:.0f} GB/s (read-only streaming reference)")


def bench_attn():
    print(f"{'case':28s} {'kv':5s} {'us':>8s} {'GB/s':>8s}")
    for name, B, H, G, hs, ctx in [("3b B=1 ctx2k", 1, 32, 32, 128, 2048), ("3b B=32 ctx2k", 32, 32, 32, 128, 2048),
                                   ("70b B=1 ctx2k", 1, 64, 8, 128, 2048), ("falcon B=1 ctx2k", 1, 71, 1, 64, 2048),
                                   ("7b B=1 ctx512", 1, 32, 32, 128, 512)]:
        for kvdt in (torch.bfloat16,):
            L = max(4, min(64, int(300e6 // (2 * B * G * ctx * hs * 2)) + 1))  # rotate over > L2 worth of cache
            kc = [torch.randn(B, G, ctx, hs, device=DEV).to(kvdt) for _ in range(L)]
            vc = [torch.randn(B, G, ctx, hs, device=DEV).to(kvdt) for _ in range(L)]
            q = torch.randn(B, H * hs, device=DEV)
            out = torch.empty(B, H * hs, device=DEV)
            pos = torch.tensor([ctx - 1], dtype=torch.int32, device=DEV)
            ws = torch.empty(lib.lp_attn_workspace_bytes(B, 1, H, hs, ctx) + 16, dtype=torch.uint8, device=DEV)
            fn = lambda i: lib.lp_attn_decode(q.data_ptr(), kc[i % L].data_ptr(), vc[i % L].data_ptr(), 1, pos.data_ptr(), out.data_ptr(),  # noqa: E731
                                              ws.data_ptr(), ws.numel(), B, 1, H, G, hs, ctx, 1 / math.sqrt(hs), 0, stream())
            us = timeit(fn, iters=30)
            nbytes = 2 * B * G * ctx * hs * 2
            print(f"{name:28s} {'bf16':5s} {us:8.2f} {nbytes / us / 1e3:8.0f}   (generic kernel, attention only)")
            qkv = torch.randn(B, (H + 2 * G) * hs, device=DEV)
            cos = torch.randn(ctx, hs, device=DEV)
            wsf = torch.zeros(lib.lp_attn_fused_workspace_bytes(B, H, G, hs, ctx) + 16, dtype=torch.uint8, device=DEV)
            fn2 = lambda i: lib.lp_attn_decode_fused(qkv.data_ptr(), cos.data_ptr(), cos.data_ptr(), pos.data_ptr(), out.data_ptr(),  # noqa: E731
                                                     kc[i % L].data_ptr(), vc[i % L].data_ptr(), 1, wsf.data_ptr(), wsf.numel(), B, H, G, hs,
                                                     hs, ctx, 1 / math.sqrt(hs), 0, stream())
            assert fn2(0) == 0
            us = timeit(fn2, iters=30)
            print(f"{name:28s} {'bf16':5s} {us:8.2f} {nbytes / us / 1e3:8.0f}   (fused rope+append+attention+merge)")


def bench_gemm():
    """lp_gemm_bf16_tc alone (prefill projections): TFLOP/s against the measured cuBLAS bf16 peak."""
    import json
    pk = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    peak = pk.get("bf16_tflops", 1672.9)
    print(f"{'shape (M,N,K)':28s} {'terms':>5s} {'us':>9s} {'TFLOP/s':>8s} {'%burst':>7s}   torch.matmul us / TF")
    shapes = [(2048, 4608, 4544), (2048, 18176, 4544), (2048, 4544, 18176), (2048, 12288, 4096), (2048, 22016, 4096), (2048, 4096, 11008),
              (4096, 8192, 8192), (32, 16384, 4096)]
    if os.environ.get("KBENCH_GEMM_SHAPES"):  # e.g. "2048x18176x4544,2048x12288x4096"
        shapes = [tuple(int(v) for v in s.split("x")) for s in os.environ["KBENCH_GEMM_SHAPES"].split(",")]
    for M, N, K in shapes:
        x = torch.randn(M, K, device=DEV)
        w = [torch.randn(N, K, device=DEV, dtype=torch.bfloat16) * 0.02 for _ in range(3)]
        out = torch.empty(M, N, device=DEV)
        for nt in (1, 2):
            terms = torch.empty(nt, M, K, dtype=torch.bfloat16, device=DEV)
            _lib.check(lib.lp_split_bf16(x.data_ptr(), terms.data_ptr(), M, K, nt, -1, None, None, 0.0, 0, stream()))
            fn = lambda i: lib.lp_gemm_bf16_tc(terms.data_ptr(), nt, M, w[i % 3].data_ptr(), N, K, None, 0, None, out.data_ptr(), None, 0, 0, stream())  # noqa: E731
            assert fn(0) == 0
            us = timeit(fn, iters=20, warm=3)
            tf = 2.0 * M * N * K / us / 1e6
            xb = x.bfloat16()
            ust = timeit(lambda i: torch.matmul(xb, w[i % 3].t()), iters=20, warm=3)
            dbg = ""
            if int(os.environ.get("LP_GEMM_DEBUG", "0")) & 2:
                import ctypes
                st = (ctypes.c_longlong * 4)()
                fn(0)
                lib.lp_debug_gemm_stats(ctypes.cast(st, ctypes.c_void_p))
                if st[2]:
                    dbg = f"   CTA0 main loops: {st[0] / st[2]:.0f} cycles/k-block, SM clock {st[0] / max(st[1], 1):.3f} GHz"
            print(f"{str((M, N, K)):28s} {nt:5d} {us:9.1f} {tf:8.1f} {100 * tf / peak:7.1f}   {ust:8.1f} / {2.0 * M * N * K / ust / 1e6:6.1f}{dbg}")


if __name__ == "__main__":
    what = sys.argv[1:] or ["linear", "attn"]
    if "linear" in what:
        bench_linear()
    if "gemm" in what:
        bench_gemm()
    if "attn" in what:
        bench_attn()
