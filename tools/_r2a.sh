#!/bin/bash
# round-2 call A: new tests first, then the whole GPU suite, traces, bench
set -x
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_decode_step_gpu.py tests/test_real_shapes_gpu.py -x -q -m gpu > gpurun_out/r2a_tests_step.log 2>&1
echo "rc=$?" >> gpurun_out/r2a_tests_step.log
tail -15 gpurun_out/r2a_tests_step.log
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2a_tests_all.log 2>&1
echo "rc=$?" >> gpurun_out/r2a_tests_all.log
tail -8 gpurun_out/r2a_tests_all.log
timeout 300 python tools/trace_decode_step.py llama2-7b-int4g128-b1 > gpurun_out/r2a_trace_int4.log 2>&1
timeout 300 python tools/trace_decode_step.py stablelm-3b-bf16-b1 > gpurun_out/r2a_trace_3b.log 2>&1
LP_DS_I4PAIR=1 timeout 300 python tools/trace_decode_step.py llama2-7b-int4g128-b1 > gpurun_out/r2a_trace_int4_pair.log 2>&1
tail -3 gpurun_out/r2a_trace_int4.log gpurun_out/r2a_trace_3b.log gpurun_out/r2a_trace_int4_pair.log
timeout 900 python bench.py --steps 64 --warmup 8 --no-cpu-baseline > gpurun_out/r2a_bench.log 2>&1
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2a_bench.log').read().strip().splitlines()[-1])
print('headline', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'])
for e in d.get('also', []):
    print({k: e[k] for k in e if k in ('workload','tok_s','ms_per_step','step_frac','prefill_tok_s','tensor_frac_of_sustained_peak','error')})
print(d.get('tp'))
PY
