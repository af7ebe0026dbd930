python -m pytest tests -m gpu -q --maxfail=30 --timeout 600 > gpurun_out/r17_tests.log 2>&1; tail -5 gpurun_out/r17_tests.log
python tools/trace_step.py llama2-7b-int4g128-b1 2>&1 | tail -9
python bench.py --steps 64 --warmup 8 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r17_bench.log; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r17_bench.log').read())
print('3b-b1 tok/s', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'frac', round(d['roofline']['whole_step']['frac'],3), 'launches', d['roofline']['whole_step']['launches'], 'kernel', d['roofline']['kernel'])
for e in d.get('also',[]): print(e['workload'], round(e.get('tok_s',0),1), 'ms', round(e.get('ms_per_step',0),3), 'frac', round(e.get('step_frac',0),3), e.get('error',''), e.get('kernel'))
PY
