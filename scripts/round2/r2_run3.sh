mkdir -p gpurun_out
export BB_LIB_PATH=$PWD/bayesic_b200/lib/libbayesic_b200_timeline.so
timeout 120 python tests/gpu_timeline.py 18944 2>&1 | tail -4
timeout 120 python tests/gpu_timeline.py 2097152 2>&1 | tail -4
unset BB_LIB_PATH
timeout 900 python tests/gpu_parity_report.py > gpurun_out/parity_default.txt 2>&1; tail -30 gpurun_out/parity_default.txt
BB_LIB_PATH=$PWD/bayesic_b200/lib/libbayesic_b200_chain4.so timeout 900 python tests/gpu_parity_report.py > gpurun_out/parity_chain4.txt 2>&1; grep "cfg[345]" gpurun_out/parity_chain4.txt | grep -v global
BB_LIB_PATH=$PWD/bayesic_b200/lib/libbayesic_b200_chain4.so timeout 600 python tests/gpu_cfg_timing.py 2>&1 | tail -12
