#!/bin/bash
export PYTHONUNBUFFERED=1
N=${1:-8}
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_tp.py -q -m gpu 2>&1 | tail -4 > gpurun_out/r2q_tests_tp_n$N.log; cat gpurun_out/r2q_tests_tp_n$N.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2q_bench_tp$N.log 2> gpurun_out/r2q_bench_tp$N.err
echo "rc=$?"; tail -2 gpurun_out/r2q_bench_tp$N.err
python - $N <<'PY'
import json, sys
d = json.loads(open(f'gpurun_out/r2q_bench_tp{sys.argv[1]}.log').read().strip().splitlines()[-1])
print('headline', round(d['value'],1), d['n_gpus'])
print('tp', d.get('tp'))
PY
