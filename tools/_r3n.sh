#!/bin/bash
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "rope or attn_prefill or attention_prefill or prefill" 2>&1 | tail -3
python tools/kbench_prefill_attn.py 2>&1 | tail -17
timeout 300 python - <<'PY'
import sys, json
sys.path.insert(0, '.')
import torch, bench
r = bench.run_prefill("falcon-7b", 1792, torch.device("cuda", 0))
print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items()})
PY
