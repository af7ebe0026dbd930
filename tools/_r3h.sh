#!/bin/bash
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "attention_decode_fused or gemm_bf16_tc" -x 2>&1 | tail -4
for i in 1 2; do
timeout 300 python bench.py --workload stablelm-3b-bf16-b32 --steps 16 --warmup 4 --no-extras --no-cpu-baseline > gpurun_out/r3h_b32.log 2>&1
python - <<'PY'
import json, sys
d = json.loads(open('gpurun_out/r3h_b32.log').read().strip().splitlines()[-1])
print('b32 tok/s', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'whole-step frac', round(d['roofline']['whole_step']['frac'],4), 'e2e', round(d['e2e']['value'],1))
PY
done
