#!/bin/bash
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_decode_step_gpu.py tests/test_real_shapes_gpu.py -x -q -m gpu -k "slab or int4" > gpurun_out/r2g_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2g_tests.log
tail -5 gpurun_out/r2g_tests.log
python tools/trace_slab.py 3 > gpurun_out/r2g_slab.log 2>&1; head -8 gpurun_out/r2g_slab.log; sed -n 19,30p gpurun_out/r2g_slab.log
timeout 300 python tools/trace_decode_step.py llama2-7b-int4g128-b1 > gpurun_out/r2g_trace_int4.log 2>&1
LP_DS_I4PAIR=1 timeout 300 python tools/trace_decode_step.py llama2-7b-int4g128-b1 > gpurun_out/r2g_trace_int4_pair.log 2>&1
cat gpurun_out/r2g_trace_int4.log | cut -c1-200
tail -n 1 gpurun_out/r2g_trace_int4_pair.log
