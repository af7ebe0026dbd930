#!/bin/bash
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests/test_gptq_quantizer.py -q -m gpu -x -s -k "llama7b_width" > gpurun_out/r3s.log 2>&1
grep -n "identical codes\|Hessian 4096\|passed\|failed" gpurun_out/r3s.log | head
