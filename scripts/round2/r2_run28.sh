mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stats.py -q -x -m gpu -k "split or weighted" 2>&1 | tail -4
BB_WP_PAIR=0 timeout 300 python -m pytest tests/test_gpu_stats.py -q -x -m gpu -k "split" 2>&1 | tail -3
timeout 1800 python -m pytest tests -q -x -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
