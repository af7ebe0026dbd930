#!/bin/bash
export PYTHONUNBUFFERED=1
timeout 600 python bench.py --steps 4 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2w_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2w_launches.csv python bench.py --steps 4 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2w_ncu_launch.log 2>&1
for w in stablelm-3b-bf16-b1 llama2-7b-int4g128-b1 falcon-7b-bf16-b1 llama2-7b-nf4-b1; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_step_kernel -s 2 -c 1 -f -o gpurun_out/r2w_step_$w python bench.py --workload $w --steps 8 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2w_ncu_$w.log 2>&1
ls -la gpurun_out/r2w_step_$w.ncu-rep
done
