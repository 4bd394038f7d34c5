mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_updates.py -q -x -m gpu > gpurun_out/pytest_updates.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_updates.log
timeout 300 python tests/gpu_cfg_timing.py loops 2>&1 | tee gpurun_out/loops_timing.log
