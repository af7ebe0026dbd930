#!/bin/bash
export PYTHONUNBUFFERED=1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r3w_tp8.log 2> gpurun_out/r3w_tp8.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r3w_tp8.log').read().strip().splitlines()[-1])
    print('N=8 headline', round(d['value'],1), 'tok/s (8 replicas); tp:', {k: (round(v,3) if isinstance(v,float) else v) for k,v in d.get('tp',{}).items() if k in ('tok_s','ms_per_step','efficiency','speedup_vs_n1','all_ranks_same_tokens','per_gpu_hbm_frac','n1_tok_s','error','n1_error')})
except Exception as e:
    print('failed', e); print(open('gpurun_out/r3w_tp8.err').read()[-1500:])
PY
