#!/bin/bash
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r2x_tests_all.log 2>&1; tail -3 gpurun_out/r2x_tests_all.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:"lp::" -c 400 --csv --log-file gpurun_out/r2x_launches.csv python bench.py --steps 4 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2x_ncu_launch.log 2>&1
grep -c "lp::" gpurun_out/r2x_launches.csv
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/r2x_bench.log 2> gpurun_out/r2x_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2x_bench.log').read().strip().splitlines()[-1])
print('headline', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],4), 'traffic', d['roofline']['traffic'], 'launches', d['gpu_launches'])
for e in d.get('also', []):
    print(json.dumps({k: e[k] for k in e if k in ('workload','tok_s','ms_per_step','step_frac','prefill_tok_s','tensor_frac_of_sustained_peak','error','e2e')})[:400])
    if 'reference_eager_cuda' in e: print('   ref cuda', e['reference_eager_cuda'].get('tok_s'))
    if 'reference_cpu' in e: print('   cfg0 ref cpu', e['reference_cpu'].get('tok_s'), 'ours', e['ours_gpu'].get('tok_s'))
print('tp', {k: d['tp'].get(k) for k in ('tok_s','ms_per_step','per_gpu_hbm_frac')})
print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['kind'], d['cpu_baseline']['cores'])
PY
