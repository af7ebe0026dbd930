#!/bin/bash
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_decode_step_gpu.py tests/test_real_shapes_gpu.py tests/test_model_gpu.py -q -m gpu -x -k "bnb" 2>&1 | tail -25
for w in llama2-7b-nf4-b1; do
timeout 600 python bench.py --steps 32 --warmup 8 --no-cpu-baseline --no-extras --workload $w > gpurun_out/r2t_bench_$w.log 2>&1
python - $w <<'PY'
import json, sys
d = json.loads(open(f'gpurun_out/r2t_bench_{sys.argv[1]}.log').read().strip().splitlines()[-1])
print(sys.argv[1], 'tok/s', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],4), 'whole', round(d['roofline']['whole_step']['frac'],4), d['roofline']['whole_step']['launches'])
PY
done
