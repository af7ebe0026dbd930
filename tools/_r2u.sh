#!/bin/bash
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_decode_step_gpu.py tests/test_real_shapes_gpu.py -q -m gpu -k "exchange or 70b" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_tp.py -q -m gpu 2>&1 | tail -2
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2u2_bench_tp2.log 2> gpurun_out/r2u2_bench_tp2.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2u2_bench_tp2.log').read().strip().splitlines()[-1])
print('tp', d.get('tp'))
PY
