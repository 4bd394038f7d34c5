mkdir -p gpurun_out
D=tests/gpu_profile_driver.py
for a in 0 1 2 3 4 5 6 7 8 11 15; do echo -n "BB_WP_ABLATE=$a  "; BB_WP_ABLATE=$a timeout 120 python $D weighted 2>&1 | tail -1; done 2>&1 | tee gpurun_out/r2_weighted_ablate.txt
