#!/bin/bash
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "gemm_bf16_tc" -x 2>&1 | tail -5
for sk in 0 1 0 1; do
LP_SWAP_STREAMK=$sk timeout 300 python bench.py --workload stablelm-3b-bf16-b32 --steps 16 --warmup 4 --no-extras --no-cpu-baseline > gpurun_out/r3f_b32_sk$sk.log 2>&1
python - $sk <<'PY'
import json, sys
try:
    d = json.loads(open(f'gpurun_out/r3f_b32_sk{sys.argv[1]}.log').read().strip().splitlines()[-1])
    print('streamk', sys.argv[1], 'tok/s', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'kernel', d['roofline'].get('kernel','')[-60:], 'frac', round(d['roofline']['frac'],3), d['roofline'].get('whole_step',{}).get('frac'))
except Exception as e:
    print('failed', e, open(f'gpurun_out/r3f_b32_sk{sys.argv[1]}.log').read()[-500:])
PY
done
