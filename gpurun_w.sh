timeout 600 python -m pytest tests/test_gpu_stats.py -q -x -m gpu -k "reject or mismatched" 2>&1 | tail -3
