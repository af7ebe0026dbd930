#!/bin/bash
export PYTHONUNBUFFERED=1
for c in e69a41c 434743d ed87960 5d0ca26 head e69a41c head; do
LP_LIB_PATH=$PWD/tools/ab/lib_$c.so timeout 300 python bench.py --workload stablelm-3b-bf16-b1 --steps 32 --warmup 8 --no-extras --no-cpu-baseline > gpurun_out/r2y_$c.log 2>&1
python - $c <<'PY'
import json, sys
try:
    d = json.loads(open(f'gpurun_out/r2y_{sys.argv[1]}.log').read().strip().splitlines()[-1])
    print(sys.argv[1], 'tok/s', round(d['value'],1), 'kernel', d['roofline']['kernel'][-22:])
except Exception as e:
    print(sys.argv[1], 'failed', open(f'gpurun_out/r2y_{sys.argv[1]}.log').read()[-300:])
PY
done
