#!/bin/bash
export PYTHONUNBUFFERED=1
for w in stablelm-3b-bf16-b1 llama2-7b-int4g128-b1 falcon-7b-bf16-b1; do
timeout 300 python bench.py --workload $w --steps 32 --warmup 8 --no-extras --no-cpu-baseline > gpurun_out/r3p_$w.log 2>&1
python - $w <<'PY'
import json, sys
d = json.loads(open(f'gpurun_out/r3p_{sys.argv[1]}.log').read().strip().splitlines()[-1])
print(sys.argv[1], 'tok/s', round(d['value'],1), 'kernel', d['roofline']['kernel'][-22:], 'frac', round(d['roofline']['frac'],4))
PY
done
