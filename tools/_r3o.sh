#!/bin/bash
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "rope" 2>&1 | tail -2
t0=$(date +%s)
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/r3o_bench.log 2> gpurun_out/r3o_bench.err
echo "bench rc=$? seconds=$(( $(date +%s) - t0 ))"
t0=$(date +%s)
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r3o_ref.log 2> gpurun_out/r3o_ref.err
echo "reference arm rc=$? seconds=$(( $(date +%s) - t0 ))"
tail -c 600 gpurun_out/r3o_ref.log
