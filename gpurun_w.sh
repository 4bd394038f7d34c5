for h in 0 1; do for pf in 0 2; do echo "hints $h prefetch $pf"; BB_FUSED_L2HINTS=$h BB_FUSED_PREFETCH=$pf timeout 100 python tests/gpu_profile_driver.py logistic; done; done
