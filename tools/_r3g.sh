#!/bin/bash
export PYTHONUNBUFFERED=1
for v in v3 v4 v3 v4; do
LP_LIB_PATH=$PWD/tools/ab/lib_$v.so timeout 300 python bench.py --workload stablelm-3b-bf16-b32 --steps 16 --warmup 4 --no-extras --no-cpu-baseline > gpurun_out/r3g_$v.log 2>&1
python - $v <<'PY'
import json, sys
try:
    d = json.loads(open(f'gpurun_out/r3g_{sys.argv[1]}.log').read().strip().splitlines()[-1])
    print(sys.argv[1], 'tok/s', round(d['value'],1), 'ms', round(d['ms_per_step'],3), d['roofline'].get('kernel','')[-45:])
except Exception as e:
    print('failed', e, open(f'gpurun_out/r3g_{sys.argv[1]}.log').read()[-500:])
PY
done
