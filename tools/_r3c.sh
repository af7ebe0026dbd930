#!/bin/bash
export PYTHONUNBUFFERED=1
for i in 1 2; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r3c_cold$i.log 2>&1
python - $i <<'PY'
import json, sys
d = json.loads(open(f'gpurun_out/r3c_cold{sys.argv[1]}.log').read().strip().splitlines()[-1])
print('run', sys.argv[1], 'tok/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'kernel', d['roofline']['kernel'][-22:], d['clocks'])
PY
done
timeout 900 python -m pytest tests/test_finetuned.py -q -m gpu -x 2>&1 | tail -30
