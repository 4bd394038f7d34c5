mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
timeout 200 python tests/gpu_cfg1_latency.py 2>&1 | tee gpurun_out/cfg1_latency.log
