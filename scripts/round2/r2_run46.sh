D=tests/gpu_profile_driver.py
timeout 300 python -m pytest tests/test_gpu_stats.py -q -x -m gpu -k "split or weighted" 2>&1 | tail -3
for rep in 1 2 3; do timeout 120 python $D weighted_split 2>&1 | tail -1; done
timeout 200 python tests/gpu_cfg_timing.py 2>&1 | grep -i "cfg3" | tail -4
