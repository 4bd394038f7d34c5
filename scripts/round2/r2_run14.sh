mkdir -p gpurun_out
D=tests/gpu_profile_driver.py
for p in 0 1 2 4 8; do echo -n "BB_WP_PREFETCH=$p  "; BB_WP_PREFETCH=$p timeout 120 python $D weighted 2>&1 | tail -1; done
for p in 0 1 2 4 8; do echo -n "BB_GRAM_PREFETCH=$p  "; BB_GRAM_PREFETCH=$p timeout 120 python $D gram 2>&1 | tail -1; done
for p in 0 1; do echo -n "BB_FUSED2_PREFETCH=$p  "; BB_FUSED2_PREFETCH=$p timeout 120 python $D logistic 2>&1 | tail -1; done
