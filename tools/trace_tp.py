"""Per-op timeline of the tensor-parallel step kernel on rank 0 (debug aid):
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/trace_tp.py [layer]"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from lit_parrot_b200 import _lib  # noqa: E402

layer = int(sys.argv[1]) if len(sys.argv) > 1 else 3
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
model, cfg = bench.build_tp_model("Llama-2-70b-hf", dev, rank, world)
model.use_cuda_graph = False
ctx = 2048
start = ctx - 40
model.kv_caches = model.build_kv_caches(torch.zeros(1, 1, device=dev), ctx)
for k, v in model.kv_caches:
    k[:, :, :start].normal_(0, 1)
    v[:, :, :start].normal_(0, 1)
lib = _lib.init(local)
per_layer = 7
nops = per_layer * cfg.n_layer + 1
trace = torch.zeros(nops, 148, 8, dtype=torch.int64, device=dev)
tok = torch.randint(0, cfg.vocab_size, (1, 1), generator=torch.Generator().manual_seed(1)).to(dev)
for i in range(5):
    dist.barrier()
    if i == 4:
        lib.lp_debug_step_trace(trace.data_ptr())
    model._forward_impl(tok, ctx, torch.tensor([start + i], device=dev), raw_logits=True)
torch.cuda.synchronize()
lib.lp_debug_step_trace(None)
dist.barrier()
if rank == 0:
    t = trace.cpu().double()
    names = ["qkv", "attn", "proj(push)", "exchange", "fc", "mlp.proj(push)", "exchange"]
    first = per_layer * layer
    t0 = t[first, :, 0].min()
    print(f"tp={world} rank 0: layers {layer}-{layer + 1}; us relative to the first CTA entering qkv({layer}); [min,max] over CTAs")
    for i in range(first, first + 2 * per_layer):
        row = t[i]

        def g(j):
            v = row[:, j]
            ok = v > 0
            return f"[{(v[ok].min() - t0) / 1e3:6.1f},{(v[ok].max() - t0) / 1e3:6.1f}]" if ok.any() else "[   -  ,   -  ]"
        print(f"{names[i % per_layer]:15s} start{g(0)} dep-met{g(1)} staged{g(2)} end{g(3)}")
    print(f"whole step {(t[nops - 1, :, 3].max() - t[0, :, 0].min()) / 1e3:.1f} us; per layer "
          f"{(t[per_layer * (cfg.n_layer - 1), :, 0].min() - t[per_layer, :, 0].min()) / 1e3 / (cfg.n_layer - 2):.1f} us")
dist.destroy_process_group()
