D=tests/gpu_profile_driver.py
T="tests/test_gpu_stats.py tests/test_gpu_passes.py"
timeout 60 python -m pytest $T -q -x -m gpu -k "regression or cfg4 or gram" 2>&1 | tail -1
BB_GRAM_SPLITS=5,9 timeout 60 python -m pytest $T -q -x -m gpu -k "regression or cfg4 or gram" 2>&1 | tail -1
for sp in 7,7 6,8 5,9 6,9; do echo -n "BB_GRAM_SPLITS=$sp  "; BB_GRAM_SPLITS=$sp timeout 40 python $D gram 2>&1 | tail -1; done
