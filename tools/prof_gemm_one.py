import os, sys, torch
sys.path.insert(0, '/root/repo')
from lit_parrot_b200 import _lib
lib = _lib.init(0)
DEV = torch.device("cuda", 0)
M, N, K = 2048, 18176, 4544
x = torch.randn(M, K, device=DEV)
w = torch.randn(N, K, device=DEV, dtype=torch.bfloat16) * 0.02
out = torch.empty(M, N, device=DEV)
terms = torch.empty(1, M, K, dtype=torch.bfloat16, device=DEV)
st = torch.cuda.current_stream().cuda_stream
lib.lp_split_bf16(x.data_ptr(), terms.data_ptr(), M, K, 1, -1, None, None, 0.0, 0, st)
for _ in range(3):
    assert lib.lp_gemm_bf16_tc(terms.data_ptr(), 1, M, w.data_ptr(), N, K, None, 0, None, out.data_ptr(), None, 0, 0, st) == 0
torch.cuda.synchronize()
