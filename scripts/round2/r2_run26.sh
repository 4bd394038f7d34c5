mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stats.py tests/test_gpu_passes.py tests/test_gpu_updates.py -q -x -m gpu -k "weighted or split or cfg3 or gmm or vmp" 2>&1 | tail -12
timeout 600 python tests/gpu_cfg_timing.py cfg3 2>&1 | tee gpurun_out/r2_cfg3_timing.txt
