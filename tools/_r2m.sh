#!/bin/bash
export PYTHONUNBUFFERED=1
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/r2m_bench.log 2> gpurun_out/r2m_bench.err
echo "rc=$?"; tail -3 gpurun_out/r2m_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2m_bench.log').read().strip().splitlines()[-1])
print('headline', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],4), 'traffic', d['roofline']['traffic'])
for e in d.get('also', []):
    print(json.dumps({k: e[k] for k in e if k not in ('bytes_per_step',)})[:700])
print('tp', d.get('tp'))
print('cpu', json.dumps(d.get('cpu_baseline'))[:600])
PY
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2m_bench_ref.log 2>&1; tail -1 gpurun_out/r2m_bench_ref.log | cut -c1-900
