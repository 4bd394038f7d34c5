mkdir -p gpurun_out
BB_SUFFSTATS_PDL=0 BB_SUFFSTATS_DYNAMIC=0 timeout 600 python -m pytest tests/test_gpu_peer.py tests/test_gpu_passes.py -q -x -m gpu -k "gaussian or cfg2" 2>&1 | tail -2
BB_WP_PAIR=0 BB_WP_EARLY=0 BB_GRAM_ABLATE=192 timeout 600 python -m pytest tests/test_gpu_stats.py -q -x -m gpu -k "split or weighted_suffstats or regression" 2>&1 | tail -2
bash scripts/gpurun_prof_one.sh logits mixture_logits_kernel r2_logits
bash scripts/gpurun_prof_one.sh softmax_split softmax_rows_split_kernel r2_softmax_split
