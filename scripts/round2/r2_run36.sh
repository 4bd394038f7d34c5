export BB_FUSED_V3=1
timeout 300 python -m pytest tests/test_gpu_stats.py -q -x -m gpu -k "logistic_reparam_stats or full_size_cfg5" 2>&1 | tail -3
for a in 0 1; do echo -n "BB_FUSED3_ABLATE=$a  "; BB_FUSED3_ABLATE=$a timeout 120 python tests/gpu_profile_driver.py logistic 2>&1 | tail -1; done
unset BB_FUSED_V3
echo -n "fused2  "; timeout 120 python tests/gpu_profile_driver.py logistic 2>&1 | tail -1
