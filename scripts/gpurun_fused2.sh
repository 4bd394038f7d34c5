mkdir -p gpurun_out
timeout 300 python tests/gpu_fused2_check.py 2>&1 | tee gpurun_out/fused2_v2.log
echo "W streamed only"; FUSED2_TIMING_ONLY=1 BB_FUSED2_W_TMEM=0 timeout 300 python tests/gpu_fused2_check.py 2>&1 | tail -1
