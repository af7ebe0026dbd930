#!/bin/bash
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_cli.py -q -m gpu -x -k "sample or generate or golden or cli" 2>&1 | tail -3
for i in 1 2; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r3r.log 2>&1
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r3r.log').read().strip().splitlines()[-1])
print('tok/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['e2e'])
PY
done
