#!/bin/bash
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,temperature.gpu,power.draw,clocks.sm,clocks.mem --format=csv,noheader
for w in stablelm-3b-bf16-b1 falcon-7b-bf16-b1 llama2-7b-int4g128-b1; do
for c in ed87960 head ed87960 head; do
LP_LIB_PATH=$PWD/tools/ab/lib_$c.so timeout 300 python bench.py --workload $w --steps 32 --warmup 8 --no-extras --no-cpu-baseline > gpurun_out/r3b_$c.log 2>&1
python - $c $w <<'PY'
import json, sys
try:
    d = json.loads(open(f'gpurun_out/r3b_{sys.argv[1]}.log').read().strip().splitlines()[-1])
    print(sys.argv[2], sys.argv[1], 'tok/s', round(d['value'],1), 'kernel', d['roofline']['kernel'][-22:])
except Exception as e:
    print(sys.argv[1], 'failed', open(f'gpurun_out/r3b_{sys.argv[1]}.log').read()[-300:])
PY
done
done
nvidia-smi --query-gpu=name,temperature.gpu,power.draw,clocks.sm,clocks.mem --format=csv,noheader
