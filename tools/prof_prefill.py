"""One Falcon-7b-shaped (4 layers) bf16 prefill, for an ncu launch list: python tools/prof_prefill.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lit_parrot_b200 as lp
dev = torch.device("cuda", 0)
cfg = lp.Config.from_name("falcon-7b", n_layer=4)
with torch.device(dev):
    torch.set_default_dtype(torch.bfloat16)
    m = lp.GPT(cfg)
    torch.set_default_dtype(torch.float32)
m.apply(m._init_weights)
m.eval().set_precision("bf16")
T = 1792
idx = torch.randint(0, cfg.vocab_size, (1, T)).to(dev)
pos = torch.arange(T, device=dev)
for _ in range(2):
    m.reset_cache()
    m._forward_impl(idx, cfg.block_size, pos, last_only=True, raw_logits=True)
torch.cuda.synchronize()
print("ok")
