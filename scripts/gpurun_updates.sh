mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_updates.py tests/test_gpu_passes.py -q -x -m gpu > gpurun_out/pytest_updates.log 2>&1; echo "pytest exit $?"; tail -25 gpurun_out/pytest_updates.log
