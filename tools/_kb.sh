timeout 900 python -m pytest tests -m gpu -q --maxfail=10 --timeout 300 > gpurun_out/r37_tests.log 2>&1; tail -4 gpurun_out/r37_tests.log
python tools/trace_step.py llama2-7b-int4g128-b1 2>&1 | tail -9 | head -5
python bench.py --steps 64 --warmup 8 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r37_bench.log; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r37_bench.log').read())
print('3b-b1 tok/s', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['whole_step']['frac'],3), 'launches', d['roofline']['whole_step']['launches'], 'kernel', d['roofline']['kernel'])
print('tp', {k:(round(v,3) if isinstance(v,float) else v) for k,v in d.get('tp',{}).items() if k!='bytes_per_step'})
for e in d.get('also',[]): print({k:(round(v,3) if isinstance(v,float) else v) for k,v in e.items() if k not in ('bytes_per_step','kernel')})
PY
