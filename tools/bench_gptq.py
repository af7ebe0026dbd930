"""Time the device GPTQ quantiser on Llama-2-7b-width blocks: `python tools/bench_gptq.py [n_layers] [n_samples] [groupsize]`.
Random-init bf16 weights, synthetic calibration tokens of block_size 2048.  The reference publishes 850 s for falcon-7b (32 layers,
128 samples) on an A100 (tutorials/quantize.md:114-117)."""
import sys
import time

import torch

sys.path.insert(0, ".")
import lit_parrot_b200 as lp  # noqa: E402
from lit_parrot_b200 import gptq  # noqa: E402

n_layers = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n_samples = int(sys.argv[2]) if len(sys.argv) > 2 else 16
groupsize = int(sys.argv[3]) if len(sys.argv) > 3 else 128
cfg = lp.Config.from_name("Llama-2-7b-hf", n_layer=n_layers, block_size=2048)
torch.manual_seed(1234)
prev = torch.get_default_dtype()
torch.set_default_dtype(torch.bfloat16)
with torch.device("cuda"):
    m = lp.GPT(cfg)
torch.set_default_dtype(prev)
m.apply(m._init_weights)
m.eval()
samples = torch.randint(0, cfg.vocab_size, (n_samples, cfg.block_size), generator=torch.Generator().manual_seed(1))
plain = m(samples[:1, :64].cuda()).float()
m.reset_cache()
torch.cuda.synchronize()
t0 = time.perf_counter()
gptq.blockwise_quantization(m, samples, "cuda", bits=4, groupsize=groupsize, batch=8, verbose=True)
torch.cuda.synchronize()
t = time.perf_counter() - t0
q = m(samples[:1, :64].cuda()).float()
print(f"layers {n_layers} samples {n_samples} groupsize {groupsize}: {t:.2f} s total (incl. lm_head {cfg.padded_vocab_size} x {cfg.n_embd}); "
      f"peak memory {torch.cuda.max_memory_allocated() / 1e9:.1f} GB; logits rms {plain.pow(2).mean().sqrt():.4f}, "
      f"quantisation rms error {(q - plain).pow(2).mean().sqrt():.4f}")
