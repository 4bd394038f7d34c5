D=tests/gpu_profile_driver.py
timeout 400 python -m pytest tests/test_gpu_stats.py tests/test_gpu_passes.py tests/test_hot_kernels_vs_reference.py tests/test_gpu_peer.py tests/test_gpu_streaming.py -q -x -m gpu -k "regression or cfg4 or gram" 2>&1 | tail -8
for rep in 1 2; do for b in 0 1; do echo -n "BB_GRAM_BALANCE=$b  "; BB_GRAM_BALANCE=$b timeout 120 python $D gram 2>&1 | tail -1; done; done
