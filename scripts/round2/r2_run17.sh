mkdir -p gpurun_out
D=tests/gpu_profile_driver.py
for rep in 1 2; do
for c in 0 1; do echo -n "BB_WP_COLLECTOR=$c  "; BB_WP_COLLECTOR=$c timeout 120 python $D weighted 2>&1 | tail -1; done
for c in 0 64; do echo -n "BB_GRAM_ABLATE=$c  "; BB_GRAM_ABLATE=$c timeout 120 python $D gram 2>&1 | tail -1; done
done
BB_WP_COLLECTOR=1 BB_GRAM_ABLATE=64 timeout 600 python -m pytest tests/test_gpu_stats.py -q -x -m gpu -k "weighted or regression" 2>&1 | tail -2
