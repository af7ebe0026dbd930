#!/bin/bash
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_cli.py tests/test_decode_step_gpu.py tests/test_real_shapes_gpu.py -q -m gpu -k "cli or chat or exchange or 70b or tp8" 2>&1 | tail -15
