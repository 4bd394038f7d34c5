D=tests/gpu_profile_driver.py
for rep in 1 2; do for c in 0 1; do echo -n "BB_LOGITS_COLLECTOR=$c  "; BB_LOGITS_COLLECTOR=$c timeout 120 python $D logits 2>&1 | tail -1; done; done
BB_LOGITS_COLLECTOR=1 timeout 300 python -m pytest tests/test_gpu_stats.py -q -x -m gpu -k "mixture_logits" 2>&1 | tail -2
