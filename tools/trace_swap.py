"""Where a launch of the swap-AB decode-batch GEMM spends its time: per-CTA globaltimer stamps of a library built with
-DLP_SWAP_TRACE (tools/ab/lib_trace.so).  `LP_LIB_PATH=tools/ab/lib_trace.so python tools/trace_swap.py [M N K epi]`."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lit_parrot_b200 import _lib  # noqa: E402

M, N, K, epi = (int(v) for v in (sys.argv[1:5] if len(sys.argv) >= 5 else (32, 16384, 4096, 0)))
lib = _lib.init(0)
raw = ctypes.CDLL(_lib.LIB_PATH)
st = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731
x = torch.randn(M, K, device="cuda")
w = [torch.randn(N, K, device="cuda", dtype=torch.bfloat16) * 0.02 for _ in range(6)]
out = torch.zeros(M, N, device="cuda")
terms = torch.empty(2, M, K, dtype=torch.bfloat16, device="cuda")
_lib.check(lib.lp_split_bf16(x.data_ptr(), terms.data_ptr(), M, K, 2, -1, None, None, 0.0, 0, st()))
res = out.data_ptr() if epi == 3 else None
for i in range(12):
    assert lib.lp_gemm_bf16_tc(terms.data_ptr(), 2, M, w[i % 6].data_ptr(), N, K, None, epi, res, out.data_ptr(), None, 0, 0, st()) == 0
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * 12288)()
assert raw.lp_debug_swap_trace(buf, 12288) == 0
h = np.array(buf, dtype=np.uint64).reshape(4, 256, 12).astype(np.float64)
used = h[0, :, 0] > 0
h = h[:, used]
order = np.argsort(np.median(h[:, :, 0], axis=1))  # the four launches in time order
h = h[order]
t0 = np.median(h[0, :, 0])
names = ["entry", "set-up done", "first loads issued", "first stage landed", "last MMA committed", "acc seen by epilogue", "epilogue done", "exit",
         "accumulator in registers", "outputs stored", "-", "-"]
print(f"M {M} N {N} K {K} epi {epi}: {h.shape[1]} CTAs; the last four launches, microseconds after the median entry of the first [min, median, max]")
for L in range(4):
    print(f" launch {L}")
    for i, nm in ((0, names[0]), (1, names[1]), (2, names[2]), (3, names[3]), (4, names[4]), (5, names[5]), (8, names[8]), (9, names[9]), (6, names[6]), (7, names[7])):
        v = (h[L, :, i] - t0) / 1e3
        print(f"  {nm:22s} {v.min():8.2f} {np.median(v):8.2f} {v.max():8.2f}")
