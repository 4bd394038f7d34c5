mkdir -p gpurun_out
bash scripts/gpurun_prof_one.sh weighted_split weighted_pairs2_kernel r2_weighted_split
bash scripts/gpurun_prof_one.sh gram gram_pair_kernel r2_gram
timeout 300 python -m pytest tests/test_gpu_passes.py -q -x -m gpu -k "cfg3" 2>&1 | tail -3
ls -la gpurun_out/prof_r2_weighted_split.ncu-rep gpurun_out/prof_r2_gram.ncu-rep
