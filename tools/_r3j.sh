#!/bin/bash
export PYTHONUNBUFFERED=1
run() { # workload lib
LP_LIB_PATH=$PWD/tools/ab/lib_$2.so timeout 300 python bench.py --workload $1 --steps 32 --warmup 8 --no-extras --no-cpu-baseline > gpurun_out/r3j.log 2>&1
python - $1 $2 <<'PY'
import json, sys
try:
    d = json.loads(open('gpurun_out/r3j.log').read().strip().splitlines()[-1])
    print(sys.argv[1], sys.argv[2], 'tok/s', round(d['value'],1), 'kernel', d['roofline']['kernel'][-22:])
except Exception as e:
    print(sys.argv[1], sys.argv[2], 'failed', open('gpurun_out/r3j.log').read()[-300:])
PY
}
run llama2-7b-int4g128-b1 cur; run llama2-7b-int4g128-b1 i4; run llama2-7b-int4g128-b1 cur; run llama2-7b-int4g128-b1 i4
run stablelm-3b-bf16-b1 cur; run stablelm-3b-bf16-b1 bf; run stablelm-3b-bf16-b1 cur; run stablelm-3b-bf16-b1 bf
run falcon-7b-bf16-b1 cur; run falcon-7b-bf16-b1 bf
