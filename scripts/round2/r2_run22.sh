mkdir -p gpurun_out
D=tests/gpu_profile_driver.py
for rep in 1 2; do
for a in 0 128; do echo -n "BB_GRAM_ABLATE=$a  "; BB_GRAM_ABLATE=$a timeout 120 python $D gram 2>&1 | tail -1; done
for a in 0 64; do echo -n "BB_FUSED2_ABLATE=$a  "; BB_FUSED2_ABLATE=$a timeout 120 python $D logistic 2>&1 | tail -1; done
done
BB_GRAM_ABLATE=128 BB_FUSED2_ABLATE=64 timeout 600 python -m pytest tests/test_gpu_stats.py -q -x -m gpu -k "regression or logistic" 2>&1 | tail -2
