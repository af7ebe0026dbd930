#!/bin/bash
export PYTHONUNBUFFERED=1
for d in 0 1 2 4; do
LP_DS_DEBUG=$d timeout 300 python tools/trace_decode_step.py llama2-7b-int4g128-b1 > gpurun_out/r2d_trace_dbg$d.log 2>&1
echo "== dbg $d"; grep -E "slab|whole" gpurun_out/r2d_trace_dbg$d.log | tail -3
done
