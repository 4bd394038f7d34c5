mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -x -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python tests/gpu_cfg_timing.py cfg3 2>&1 | tee gpurun_out/r2_cfg3_timing.txt
