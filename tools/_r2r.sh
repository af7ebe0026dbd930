#!/bin/bash
export PYTHONUNBUFFERED=1
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/trace_tp.py 3 > gpurun_out/r2r_trace_tp$N.log 2>&1
grep -v "OMP_NUM\|\*\*\*" gpurun_out/r2r_trace_tp$N.log | tail -18
