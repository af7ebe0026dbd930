#!/bin/bash
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests/test_decode_step_gpu.py tests/test_real_shapes_gpu.py tests/test_model_gpu.py -q -m gpu -x 2>&1 | tail -6
timeout 300 python tools/trace_decode_step.py llama2-7b-int4g128-b1 2>&1 | tail -7 | cut -c1-200
for w in stablelm-3b-bf16-b1 llama2-7b-int4g128-b1 falcon-7b-bf16-b1; do
timeout 300 python bench.py --workload $w --steps 32 --warmup 8 --no-extras --no-cpu-baseline > gpurun_out/r3a_$w.log 2>&1
python - $w <<'PY'
import json, sys
d = json.loads(open(f'gpurun_out/r3a_{sys.argv[1]}.log').read().strip().splitlines()[-1])
print(sys.argv[1], 'tok/s', round(d['value'],1), 'kernel', d['roofline']['kernel'][-22:], 'frac', round(d['roofline']['frac'],4))
PY
done
