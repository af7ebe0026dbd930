#!/bin/bash
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_decode_step_gpu.py -x -q -m gpu -k "slab" 2>&1 | tail -3
for d in 0 1; do
LP_DS_DEBUG=$d python tools/trace_slab.py 3 > gpurun_out/r2k_slab_$d.log 2>&1
echo "== dbg $d"; sed -n 9,10p gpurun_out/r2k_slab_$d.log; sed -n 19,24p gpurun_out/r2k_slab_$d.log
done
timeout 300 python tools/trace_decode_step.py llama2-7b-int4g128-b1 2>&1 | tail -1
