#!/bin/bash
export PYTHONUNBUFFERED=1
N=8
timeout 600 python -m pytest tests/test_tp.py -q -m gpu -k "oracle and 8" 2>&1 | tail -2 | tee gpurun_out/r2v_tests_tp8.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2v_bench_tp$N.log 2> gpurun_out/r2v_bench_tp$N.err
python - $N <<'PY'
import json, sys
d = json.loads(open(f'gpurun_out/r2v_bench_tp{sys.argv[1]}.log').read().strip().splitlines()[-1])
print('headline', round(d['value'],1), d['n_gpus'])
print('tp', d.get('tp'))
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/trace_tp.py 3 > gpurun_out/r2v_trace_tp$N.log 2>&1
grep -v "OMP_NUM\|\*\*\*" gpurun_out/r2v_trace_tp$N.log | tail -9
