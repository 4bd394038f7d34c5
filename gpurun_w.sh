mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stats.py tests/test_gpu_passes.py -q -x -m gpu -k "mixture_logits or cfg3" > gpurun_out/pytest_w.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/pytest_w.log
timeout 100 python tests/gpu_profile_driver.py logits
