mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_passes.py tests/test_gpu_stats.py -q -m gpu -k "cfg3 or weighted or mixture" > gpurun_out/pytest_w.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_w.log
