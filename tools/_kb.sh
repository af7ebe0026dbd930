python -m pytest tests -m gpu -q --maxfail=30 --timeout 600 > gpurun_out/r9_tests.log 2>&1; tail -8 gpurun_out/r9_tests.log
for kb in 112 160 216; do echo "== budget $kb"; LP_GS_BUDGET_KB=$kb python bench.py --steps 64 --warmup 8 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('3b-b1 tok/s', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'frac', round(d['roofline']['whole_step']['frac'],3), 'launches', d['roofline']['whole_step']['launches'], 'kernel', d['roofline']['kernel'])
for e in d.get('also',[]): print(e['workload'], round(e.get('tok_s',0),1), 'ms', round(e.get('ms_per_step',0),3), 'frac', round(e.get('step_frac',0),3), e.get('error',''))
"; done > gpurun_out/r9_bench.log 2>&1
cat gpurun_out/r9_bench.log
