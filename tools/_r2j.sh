#!/bin/bash
export PYTHONUNBUFFERED=1
for d in 0 8 16 24 32 64 96; do
LP_DS_DEBUG=$d python tools/trace_slab.py 3 > gpurun_out/r2j_slab_$d.log 2>&1
echo "== dbg $d"; sed -n 9,10p gpurun_out/r2j_slab_$d.log; sed -n 19,22p gpurun_out/r2j_slab_$d.log
done
