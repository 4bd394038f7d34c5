for h in 0 1 2 4 8; do echo "gram prefetch $h"; BB_GRAM_PREFETCH=$h timeout 100 python tests/gpu_profile_driver.py gram; done
