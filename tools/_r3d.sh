#!/bin/bash
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gptq_quantizer.py -q -m gpu -s --tb=short -k "blockwise or cli" > gpurun_out/r3d_gptq.log 2>&1
grep -n "identical codes\|passed\|failed\|Error\|error:" gpurun_out/r3d_gptq.log | head -40
