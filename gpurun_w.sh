mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stats.py tests/test_gpu_passes.py -q -x -m gpu -k "logistic or cfg5" > gpurun_out/pytest_w.log 2>&1; echo "pytest exit $?"; tail -12 gpurun_out/pytest_w.log
timeout 200 python tests/gpu_cfg_timing.py cfg5 2>&1 | tail -2
BB_LOGISTIC_UNFUSED=1 timeout 200 python tests/gpu_cfg_timing.py cfg5 2>&1 | tail -1
