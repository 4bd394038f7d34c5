mkdir -p gpurun_out
D=tests/gpu_profile_driver.py
for rep in 1 2; do
echo -n "baseline  "; timeout 120 python $D weighted 2>&1 | tail -1
echo -n "BB_WP_EARLY=1  "; BB_WP_EARLY=1 timeout 120 python $D weighted 2>&1 | tail -1
echo -n "BB_WP_EARLY=1 BB_WP_L1=1 BB_WP_PREFETCH=2  "; BB_WP_EARLY=1 BB_WP_L1=1 BB_WP_PREFETCH=2 timeout 120 python $D weighted 2>&1 | tail -1
done
BB_WP_EARLY=1 timeout 600 python -m pytest tests/test_gpu_stats.py -q -x -m gpu -k "weighted" 2>&1 | tail -2
