#!/bin/bash
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r2l_tests_all.log 2>&1
echo "rc=$?" >> gpurun_out/r2l_tests_all.log
tail -6 gpurun_out/r2l_tests_all.log
# the test that hung once in round 1, looped, default build and with the paired int4 loop (watchdog armed: a hang becomes a failure)
timeout 900 python -m pytest "tests/test_decode_step_gpu.py::test_step_kernel_greedy_tokens_and_sliding_window" -q -m gpu --count 1 -p no:cacheprovider > /dev/null 2>&1
for i in $(seq 1 25); do timeout 120 python -m pytest "tests/test_decode_step_gpu.py::test_step_kernel_greedy_tokens_and_sliding_window" -q -m gpu -x 2>&1 | tail -1; done > gpurun_out/r2l_loop_default.log 2>&1
sort gpurun_out/r2l_loop_default.log | uniq -c
for i in $(seq 1 25); do LP_DS_I4PAIR=1 timeout 120 python -m pytest "tests/test_decode_step_gpu.py::test_step_kernel_greedy_tokens_and_sliding_window" -q -m gpu -x 2>&1 | tail -1; done > gpurun_out/r2l_loop_pair.log 2>&1
sort gpurun_out/r2l_loop_pair.log | uniq -c
# synccheck / racecheck on the step kernel (small model; watchdog off: the tools slow the kernel down by orders of magnitude)
LP_DS_TIMEOUT_MS=0 timeout 900 compute-sanitizer --tool synccheck python -m pytest "tests/test_decode_step_gpu.py::test_step_kernel_logits_int4" -q -m gpu -x > gpurun_out/r2l_synccheck.log 2>&1
tail -4 gpurun_out/r2l_synccheck.log
LP_DS_TIMEOUT_MS=0 timeout 1200 compute-sanitizer --tool racecheck python -m pytest "tests/test_decode_step_gpu.py::test_step_kernel_logits_int4" -q -m gpu -x > gpurun_out/r2l_racecheck.log 2>&1
tail -4 gpurun_out/r2l_racecheck.log
