mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -x -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python tests/gpu_cfg1_latency.py 2>&1 | tee gpurun_out/r2_cfg1_latency.txt
timeout 900 python tests/gpu_parity_report.py > gpurun_out/r2_parity_report.txt 2>&1; tail -32 gpurun_out/r2_parity_report.txt
