mkdir -p gpurun_out
T=tests/cuda/gram_test
timeout 200 $T 2>&1 | tail -5
for a in 0 1 33; do echo "ablate=$a"; BB_GRAM_ABLATE=$a timeout 60 $T 1048576 1024 5 2>&1 | tail -1; done | tee gpurun_out/gram_ablate.log
