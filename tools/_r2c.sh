#!/bin/bash
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_decode_step_gpu.py tests/test_real_shapes_gpu.py -x -q -m gpu > gpurun_out/r2c_tests_step.log 2>&1
echo "rc=$?" >> gpurun_out/r2c_tests_step.log
tail -25 gpurun_out/r2c_tests_step.log
timeout 300 python tools/trace_decode_step.py llama2-7b-int4g128-b1 > gpurun_out/r2c_trace_int4.log 2>&1
LP_DS_FUSE=0 timeout 300 python tools/trace_decode_step.py llama2-7b-int4g128-b1 > gpurun_out/r2c_trace_int4_nofuse.log 2>&1
LP_DS_I4PAIR=1 timeout 300 python tools/trace_decode_step.py llama2-7b-int4g128-b1 > gpurun_out/r2c_trace_int4_pair.log 2>&1
cat gpurun_out/r2c_trace_int4.log
tail -n 2 gpurun_out/r2c_trace_int4_nofuse.log gpurun_out/r2c_trace_int4_pair.log
for w in llama2-7b-int4g128-b1; do
timeout 600 python bench.py --steps 64 --warmup 8 --no-cpu-baseline --no-extras --workload $w > gpurun_out/r2c_bench_$w.log 2>&1
python - $w <<'PY'
import json, sys
d = json.loads(open(f'gpurun_out/r2c_bench_{sys.argv[1]}.log').read().strip().splitlines()[-1])
print(sys.argv[1], 'tok/s', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],4), 'whole', round(d['roofline']['whole_step']['frac'],4), d['roofline']['whole_step']['bytes'])
PY
done
