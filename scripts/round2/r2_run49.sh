D=tests/gpu_profile_driver.py
timeout 300 python -m pytest tests/test_gpu_stats.py tests/test_gpu_passes.py -q -x -m gpu -k "regression or cfg4 or gram" 2>&1 | tail -1
for pf in 0 16 8 32 64 0 16; do echo -n "BB_GRAM_PREFETCH=$pf  "; BB_GRAM_PREFETCH=$pf timeout 120 python $D gram 2>&1 | tail -1; done
