mkdir -p gpurun_out
D=tests/gpu_profile_driver.py
BB_LIB_PATH=$PWD/bayesic_b200/lib/libbayesic_b200_wptl.so timeout 120 python $D weighted 2>&1 | tail -6 | tee gpurun_out/r2_weighted_timeline.txt
echo -n "gram default  "; timeout 120 python $D gram 2>&1 | tail -1
