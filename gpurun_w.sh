mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stats.py -q -x -m gpu -k "full_size or large_cfg3" --durations=5 > gpurun_out/pytest_w.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_w.log
