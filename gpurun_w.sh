mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_plan_parity.py tests/test_gpu_passes.py -q -x -m gpu > gpurun_out/pytest_w.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/pytest_w.log
timeout 300 python tests/gpu_cfg_timing.py cfg4 cfg5 2>&1 | tee gpurun_out/cfg45_timing.log
