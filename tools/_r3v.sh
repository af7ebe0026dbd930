#!/bin/bash
export PYTHONUNBUFFERED=1
for w in stablelm-3b-bf16-b1 falcon-7b-bf16-b1; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_step_kernel -s 2 -c 1 -f -o gpurun_out/r3v_step_$w python bench.py --workload $w --steps 8 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r3v_ncu_$w.log 2>&1
ls -la gpurun_out/r3v_step_$w.ncu-rep
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3v_launches.csv python bench.py --steps 4 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r3v_ncu_launch.log 2>&1
wc -l gpurun_out/r3v_launches.csv
