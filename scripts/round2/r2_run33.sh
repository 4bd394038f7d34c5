mkdir -p gpurun_out
export BB_FUSED_V3=1
timeout 300 python -m pytest tests/test_gpu_stats.py -q -x -m gpu -k "logistic_reparam_stats or full_size_cfg5" 2>&1 | tail -15
timeout 120 python tests/gpu_profile_driver.py logistic 2>&1 | tail -1
unset BB_FUSED_V3
timeout 120 python tests/gpu_profile_driver.py logistic 2>&1 | tail -1
