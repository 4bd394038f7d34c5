mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_stats.py -q -x -m gpu -k "split" 2>&1 | tail -12
BB_WP_PAIR=0 timeout 300 python -m pytest tests/test_gpu_stats.py -q -x -m gpu -k "split" 2>&1 | tail -3
timeout 300 python -m pytest tests/test_gpu_passes.py tests/test_gpu_updates.py -q -x -m gpu -k "cfg3 or gmm or vmp" 2>&1 | tail -3
timeout 300 python tests/gpu_cfg_timing.py cfg3 2>&1 | tee gpurun_out/r2_cfg3_timing.txt
