"""Phase timeline of the streaming GEMV (debug aid): python tools/trace_stream.py"""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lit_parrot_b200 import _lib
from lit_parrot_b200._lib import LpWeight
from lit_parrot_b200.quantize import tile_major_aux
DEV = torch.device("cuda", 0)
lib = _lib.init(0)
st = lambda: torch.cuda.current_stream().cuda_stream
def run(fmt, N, K, norm):
    x = torch.randn(1, K, device=DEV); out = torch.empty(1, N, device=DEV)
    nw = torch.ones(K, device=DEV)
    recs, keep = [], []
    for i in range(4):
        if fmt == "bf16":
            w = torch.randn(N, K, device=DEV, dtype=torch.bfloat16) * 0.02
            keep.append(w); recs.append(LpWeight(w.data_ptr(), None, None, None, None, _lib.LP_W_BF16, N, K, 0, 0, 0))
        else:
            rb = lib.lp_int4_row_bytes(K)
            w = torch.randint(0, 256, (N, rb), device=DEV, dtype=torch.uint8)
            sc = (torch.rand(N, K // 128, device=DEV) * 0.01).bfloat16().float(); ze = torch.full((N, K // 128), 8.0, device=DEV)
            aux2, flags = tile_major_aux(sc, ze); keep += [w, sc, ze, aux2]
            recs.append(LpWeight(w.data_ptr(), sc.data_ptr(), ze.data_ptr(), aux2.data_ptr(), None, _lib.LP_W_INT4, N, K, 128, flags, 0))
    tr = torch.zeros(4, 148 * 8, dtype=torch.int64, device=DEV)
    lib.lp_set_linear_path(2)
    def call(i):
        if norm:
            return lib.lp_norm_linear(1, nw.data_ptr(), None, 1e-5, x.data_ptr(), 1, ctypes.byref(recs[i]), 0, None, out.data_ptr(), 0, st())
        return lib.lp_linear(x.data_ptr(), 1, ctypes.byref(recs[i]), 0, None, out.data_ptr(), 0, st())
    for i in range(4): assert call(i) == 0
    torch.cuda.synchronize()
    for i in range(4):
        lib.lp_debug_stream_trace(tr[i].data_ptr()); assert call(i) == 0
    torch.cuda.synchronize(); lib.lp_debug_stream_trace(None)
    t = tr.cpu().view(4, 148, 8).double()
    t0 = t[0, :, 0].min()
    names = ["start", "waited", "staged", "stage0", "stage1", "stage2", "stage3", "end"]
    print(f"--- {fmt} N={N} K={K} norm={norm}: 4 back-to-back launches; times in us relative to first CTA start of launch 0")
    for i in range(4):
        row = t[i]
        ok = row[:, 7] > 0
        print(f"launch {i}: " + "  ".join(f"{n}:[{(row[ok, j].min() - t0) / 1e3:6.1f},{(row[ok, j].max() - t0) / 1e3:6.1f}]" for j, n in enumerate(names)))
for fmt, N, K in (("int4", 22016, 4096), ("int4", 4096, 4096), ("bf16", 16384, 4096), ("bf16", 4096, 4096)):
    run(fmt, N, K, False)
run("int4", 22016, 4096, True)
