timeout 900 python -m pytest tests -m gpu -q --maxfail=10 --timeout 300 -k "gemm_bf16_tc or tiny_models or real_width" > gpurun_out/r26_tests.log 2>&1; tail -5 gpurun_out/r26_tests.log
python - <<'PY'
import sys, torch, json
sys.path.insert(0, '.')
import bench
print(json.dumps(bench.run_prefill("falcon-7b", 1792, torch.device("cuda", 0))))
PY
