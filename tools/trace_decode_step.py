"""Per-op timeline of the persistent decode-step kernel (debug aid): python tools/trace_decode_step.py [workload] [layer]

For every op of two consecutive layers: when the CTAs start it, when its dependency is met, when the activations are staged
and when the CTAs finish it ([min, max] over CTAs, us relative to the first op shown), from lp_debug_step_trace."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from lit_parrot_b200 import _lib  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "llama2-7b-int4g128-b1"
layer = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
model, cfg, B, ctx = bench.build_model(wl, dev)
model.use_cuda_graph = False
start = ctx - 40
model.kv_caches = model.build_kv_caches(torch.zeros(B, 1, device=dev), ctx)
for k, v in model.kv_caches:
    k[:, :, :start].normal_(0, 1)
    v[:, :, :start].normal_(0, 1)
lib = _lib.init(0)
nops = 5 * cfg.n_layer + 1
trace = torch.zeros(nops, 148, 8, dtype=torch.int64, device=dev)
tok = torch.randint(0, cfg.vocab_size, (1, 1), device=dev)
for i in range(4):
    if i == 3:
        lib.lp_debug_step_trace(trace.data_ptr())
    model._forward_impl(tok, ctx, torch.tensor([start + i], device=dev), raw_logits=True)
torch.cuda.synchronize()
lib.lp_debug_step_trace(None)
t = trace.cpu().double()
par = cfg.parallel_residual
names = ["qkv", "fc", "attn", "proj", "mlp.proj"] if par else ["qkv", "attn", "proj", "fc", "mlp.proj"]
eng = model._get_engine(dev)
if eng._slabs:
    names = ["qkv", "attn", "proj*slab", "fc", "mlp*slab"]
import ctypes  # noqa: E402
ent = next(v for v in eng._steps.values() if v is not None)
print("cooperative launch:", lib.lp_decode_step_cooperative(ctypes.byref(ent[0])), " fused slabs:", bool(eng._slabs))
first = 5 * layer
t0 = t[first, :, 0].min()
print(f"{wl}: layers {layer}-{layer + 1}; us relative to the first CTA entering {names[0]}({layer}); [min,max] over CTAs")
for i in range(first, min(first + 10, nops)):
    row = t[i]
    f = lambda j: f"[{(row[:, j].min() - t0) / 1e3:6.1f},{(row[:, j].max() - t0) / 1e3:6.1f}]"  # noqa: E731
    staged = row[:, 2]
    ok = staged > 0
    sg = f"[{(staged[ok].min() - t0) / 1e3:6.1f},{(staged[ok].max() - t0) / 1e3:6.1f}]" if ok.any() else "[   -  ,   -  ]"
    name = names[i % 5] if i < nops - 1 else "lm_head"
    def g(j):
        v = row[:, j]
        okj = v > 0
        return f"[{(v[okj].min() - t0) / 1e3:6.1f},{(v[okj].max() - t0) / 1e3:6.1f}]" if okj.any() else "[   -  ,   -  ]"
    print(f"{name:9s} start{f(0)} dep-met{f(1)} xload{g(4)} norm{g(5)} amax{g(6)} staged{sg} end{f(3)}")
tot = (t[nops - 1, :, 3].max() - t[0, :, 0].min()) / 1e3
print(f"whole step kernel: {tot:.1f} us; per layer {(t[5 * (cfg.n_layer - 1), :, 0].min() - t[5, :, 0].min()) / 1e3 / (cfg.n_layer - 2):.1f} us")
