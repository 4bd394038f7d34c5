mkdir -p gpurun_out
D=tests/gpu_profile_driver.py
for p in 0 1 2 4 8 16; do echo -n "BB_WP_PREFETCH=$p  "; BB_WP_PREFETCH=$p timeout 120 python $D weighted 2>&1 | tail -1; done
echo -n "gram default (collector on)  "; timeout 120 python $D gram 2>&1 | tail -1
BB_WP_PREFETCH=4 timeout 600 python -m pytest tests/test_gpu_stats.py -q -x -m gpu -k "weighted or regression" 2>&1 | tail -2
