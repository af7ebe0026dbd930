#!/bin/bash
export PYTHONUNBUFFERED=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_step_kernel -s 2 -c 1 -f -o gpurun_out/r2h_int4 python bench.py --workload llama2-7b-int4g128-b1 --steps 8 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2h_ncu_int4.log 2>&1
tail -3 gpurun_out/r2h_ncu_int4.log | cut -c1-300
ls -la gpurun_out/r2h_int4.ncu-rep
