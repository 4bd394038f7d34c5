mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_passes.py tests/test_gpu_stats.py -q -x -m gpu -k "cfg3 or weighted" > gpurun_out/pytest_w.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/pytest_w.log
timeout 200 python tests/gpu_cfg_timing.py cfg3 2>&1 | tail -6
