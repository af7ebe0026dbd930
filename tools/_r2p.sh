#!/bin/bash
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_tp.py -q -m gpu -x 2>&1 | grep -v "^$" | tail -60
