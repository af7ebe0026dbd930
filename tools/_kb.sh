timeout 900 python -m pytest tests -m gpu -q --maxfail=10 --timeout 300 > gpurun_out/r24_tests.log 2>&1; tail -12 gpurun_out/r24_tests.log
