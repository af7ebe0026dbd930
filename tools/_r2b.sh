#!/bin/bash
set -x
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_decode_step_gpu.py tests/test_real_shapes_gpu.py tests/test_kernels_gpu.py -q -m gpu -k "step or real or shapes or sample or widths" > gpurun_out/r2b_tests_step.log 2>&1
echo "rc=$?" >> gpurun_out/r2b_tests_step.log
tail -12 gpurun_out/r2b_tests_step.log
timeout 300 python tools/trace_decode_step.py llama2-7b-int4g128-b1 > gpurun_out/r2b_trace_int4.log 2>&1
timeout 300 python tools/trace_decode_step.py stablelm-3b-bf16-b1 > gpurun_out/r2b_trace_3b.log 2>&1
LP_DS_I4PAIR=1 timeout 300 python tools/trace_decode_step.py llama2-7b-int4g128-b1 > gpurun_out/r2b_trace_int4_pair.log 2>&1
LP_DS_COOP=0 timeout 300 python tools/trace_decode_step.py llama2-7b-int4g128-b1 > gpurun_out/r2b_trace_int4_nocoop.log 2>&1
for f in gpurun_out/r2b_trace_int4.log gpurun_out/r2b_trace_3b.log gpurun_out/r2b_trace_int4_pair.log gpurun_out/r2b_trace_int4_nocoop.log; do tail -n 2 $f; done
for w in stablelm-3b-bf16-b1 llama2-7b-int4g128-b1 falcon-7b-bf16-b1; do
timeout 600 python bench.py --steps 64 --warmup 8 --no-cpu-baseline --no-extras --workload $w > gpurun_out/r2b_bench_$w.log 2>&1
python - $w <<'PY'
import json, sys
d = json.loads(open(f'gpurun_out/r2b_bench_{sys.argv[1]}.log').read().strip().splitlines()[-1])
print(sys.argv[1], 'tok/s', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],4), 'whole', round(d['roofline']['whole_step']['frac'],4))
PY
done
