#!/bin/bash
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r3k_tests.log 2>&1; tail -5 gpurun_out/r3k_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_swap_kernel -s 6 -c 1 -f -o gpurun_out/r3k_swap_m32 env KBENCH_M=32 KBENCH_SHAPES="16384,4096,1" python tools/kbench_skinny.py > gpurun_out/r3k_ncu_swap.log 2>&1
ls -la gpurun_out/r3k_swap_m32.ncu-rep
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r3k_b32_launches.csv python bench.py --workload stablelm-3b-bf16-b32 --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r3k_ncu_b32.log 2>&1
wc -l gpurun_out/r3k_b32_launches.csv
