export BB_FUSED_V3=1
for a in 0 1 2 4 8 12 16 18 32 33 63; do echo -n "BB_FUSED3_ABLATE=$a  "; BB_FUSED3_ABLATE=$a timeout 120 python tests/gpu_profile_driver.py logistic 2>&1 | tail -1; done
