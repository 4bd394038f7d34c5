mkdir -p gpurun_out
D=tests/gpu_profile_driver.py
echo -n "baseline  "; timeout 120 python $D weighted 2>&1 | tail -1
for p in 0 1 2 3 4; do echo -n "BB_WP_L1=1 BB_WP_PREFETCH=$p  "; BB_WP_L1=1 BB_WP_PREFETCH=$p timeout 120 python $D weighted 2>&1 | tail -1; done
BB_WP_L1=1 BB_WP_PREFETCH=2 timeout 600 python -m pytest tests/test_gpu_stats.py -q -x -m gpu -k "weighted" 2>&1 | tail -2
