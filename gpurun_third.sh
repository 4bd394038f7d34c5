mkdir -p gpurun_out
timeout 600 python tests/gpu_first_light.py > gpurun_out/first_light.log 2>&1; echo "first_light exit $?"; head -17 gpurun_out/first_light.log
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/pytest_gpu.log
