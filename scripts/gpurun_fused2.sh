mkdir -p gpurun_out
timeout 60 tests/cuda/mma_ss_rate 2>&1 | tee gpurun_out/mma_ss_rate.log
export BB_FUSED_V2=1 FUSED2_TIMING_ONLY=1
echo base; timeout 300 python tests/gpu_fused2_check.py 2>&1 | tail -1
echo interleave_b; BB_FUSED2_INTERLEAVE_B=1 timeout 300 python tests/gpu_fused2_check.py 2>&1 | tail -1
