#!/bin/bash
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "gemm_bf16_tc" -x 2>&1 | tail -4
KBENCH_M=32 python tools/kbench_skinny.py 2>&1 | tail -6
LP_LIB_PATH=$PWD/tools/ab/lib_trace.so python tools/trace_swap.py 32 16384 4096 1 2>&1 | tail -24 | head -12
timeout 300 python bench.py --workload stablelm-3b-bf16-b32 --steps 16 --warmup 4 --no-extras --no-cpu-baseline > gpurun_out/r3i_b32.log 2>&1
python - <<'PY'
import json, sys
d = json.loads(open('gpurun_out/r3i_b32.log').read().strip().splitlines()[-1])
print('b32 tok/s', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'whole-step frac', round(d['roofline']['whole_step']['frac'],4), 'e2e', round(d['e2e']['value'],1), d['roofline']['kernel'][-50:])
PY
