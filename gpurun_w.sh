mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stats.py -q -x -m gpu -k "weighted" > gpurun_out/pytest_w.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_w.log
timeout 300 python tests/gpu_weighted_timing.py 2>&1 | tee gpurun_out/weighted_timing.log
