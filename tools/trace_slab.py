"""Debug aid: per-CTA breakdown of the SLAB ops of one layer (python tools/trace_slab.py [layer])."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from lit_parrot_b200 import _lib  # noqa: E402

layer = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda", 0)
model, cfg, B, ctx = bench.build_model("llama2-7b-int4g128-b1", dev)
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
t = trace.cpu()
for name, off in (("proj*slab", 2), ("mlp*slab", 4)):
    r = t[5 * layer + off].double()
    print(name, "cta: start staged first-stage loop-end end | wait-us nunits nseg")
    t0 = r[:, 0].min()
    for c in list(range(0, 148, 9)):
        meta = int(t[5 * layer + off][c, 7])
        print(f"  {c:3d}: {(r[c,0]-t0)/1e3:6.1f} {(r[c,2]-t0)/1e3:6.1f} {(r[c,4]-t0)/1e3:6.1f} {(r[c,5]-t0)/1e3:6.1f} {(r[c,3]-t0)/1e3:6.1f} | "
              f"{r[c,6]/1965.0:6.2f} {meta >> 8} {meta & 255}")
