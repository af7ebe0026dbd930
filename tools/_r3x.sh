#!/bin/bash
export PYTHONUNBUFFERED=1
for w in stablelm-3b-bf16-b1 llama2-7b-int4g128-b1 falcon-7b-bf16-b1; do
for v in 0 1; do
LP_DS_QKV_SK=$v timeout 300 python bench.py --workload $w --steps 32 --warmup 8 --no-extras --no-cpu-baseline > gpurun_out/r3x.log 2>&1
python - $w $v <<'PY'
import json, sys
try:
    d = json.loads(open('gpurun_out/r3x.log').read().strip().splitlines()[-1])
    print(sys.argv[1], 'qkv_sk', sys.argv[2], 'tok/s', round(d['value'],1), 'kernel', d['roofline']['kernel'][-22:], 'frac', round(d['roofline']['frac'],4))
except Exception as e:
    print(sys.argv[1], sys.argv[2], 'failed', open('gpurun_out/r3x.log').read()[-400:])
PY
done
done
timeout 900 python -m pytest tests/test_decode_step_gpu.py tests/test_real_shapes_gpu.py tests/test_model_gpu.py -q -m gpu -x 2>&1 | tail -3
