mkdir -p gpurun_out
timeout 300 python tests/gpu_fused2_check.py 2>&1 | tee gpurun_out/fused2_v2.log
BB_FUSED_V2=0 timeout 300 python tests/gpu_fused2_check.py 2>&1 | tee gpurun_out/fused2_v1.log
timeout 60 tests/cuda/mma_ss_rate 2>&1 | tee gpurun_out/mma_ss_rate.log
