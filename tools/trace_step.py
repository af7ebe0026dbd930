"""Timeline of the streaming GEMVs inside one replayed decode step (debug aid): python tools/trace_step.py [workload]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from lit_parrot_b200 import _lib
wl = sys.argv[1] if len(sys.argv) > 1 else "llama2-7b-int4g128-b1"
dev = torch.device("cuda", 0)
model, cfg, B, ctx = bench.build_model(wl, dev)
start = ctx - 40
model.kv_caches = model.build_kv_caches(torch.zeros(B, 1, device=dev), ctx)
tok0 = torch.randint(0, cfg.vocab_size, (B, 1))
eng = model._get_engine(dev)
model.rope_cache = model.build_rope_cache(tok0.to(dev)); eng.set_rope(model.rope_cache)
ncalls = 4 * cfg.n_layer
eng.trace = torch.zeros(ncalls, 148 * 8, dtype=torch.int64, device=dev)
st = eng.gen_state(cfg.block_size)
st["seq"].zero_(); st["pos"].fill_(start)
replay = eng.decode_step(model.kv_caches, 1.0, 1, 1234, B)
for _ in range(5): replay()
torch.cuda.synchronize()
t = eng.trace.cpu().view(ncalls, 148, 8).double()
names = ["qkv", "proj", "fc", "mlp.proj"]
t0 = t[4 * 3, :, 0].min()
print(f"{wl}: layers 3-4, us relative to qkv(3) first CTA start; [min,max] over CTAs")
for i in range(4 * 3, 4 * 5):
    row = t[i]; ok = row[:, 7] > 0
    f = lambda j: f"[{(row[ok, j].min() - t0) / 1e3:6.1f},{(row[ok, j].max() - t0) / 1e3:6.1f}]"
    print(f"{names[i % 4]:9s} ctas={int(ok.sum()):3d} start{f(0)} waited{f(1)} staged{f(2)} stage0{f(3)} end{f(7)}")
