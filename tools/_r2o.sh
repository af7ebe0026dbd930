#!/bin/bash
export PYTHONUNBUFFERED=1
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_tp.py -q -m gpu -x 2>&1 | tail -8
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2o_bench_tp2.log 2> gpurun_out/r2o_bench_tp2.err
echo "rc=$?"; tail -3 gpurun_out/r2o_bench_tp2.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2o_bench_tp2.log').read().strip().splitlines()[-1])
print('headline', round(d['value'],1), d['n_gpus'])
print('tp', d.get('tp'))
PY
