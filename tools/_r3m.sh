#!/bin/bash
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_tp.py -q -m gpu 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 16 --warmup 4 > gpurun_out/r3m_tp2.log 2> gpurun_out/r3m_tp2.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r3m_tp2.log').read().strip().splitlines()[-1])
    print('N=2 headline', round(d['value'],1), 'tok/s; tp:', {k: (round(v,3) if isinstance(v,float) else v) for k,v in d.get('tp',{}).items() if k in ('tok_s','ms_per_step','efficiency','speedup_vs_n1','all_ranks_same_tokens','per_gpu_hbm_frac','n1_tok_s','error')})
except Exception as e:
    print('failed', e); print(open('gpurun_out/r3m_tp2.err').read()[-1500:])
PY
