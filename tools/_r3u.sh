#!/bin/bash
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r3u_tests.log 2>&1; tail -3 gpurun_out/r3u_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/r3u_bench.log 2> gpurun_out/r3u_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3u_bench.log').read().strip().splitlines()[-1])
print('headline', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],4), 'whole', round(d['roofline']['whole_step']['frac'],4), d['clocks'])
for a in d.get('also',[]):
    print(' ', a.get('workload','')[:60], {k:(round(v,1) if isinstance(v,float) else v) for k,v in a.items() if k in ('tok_s','ms_per_step','step_frac','prefill_tok_s','tensor_frac_of_sustained_peak','ms','error')}, (a.get('e2e') or {}).get('value'))
print('tp', {k:(round(v,3) if isinstance(v,float) else v) for k,v in d.get('tp',{}).items() if k in ('tok_s','per_gpu_hbm_frac','error')})
PY
