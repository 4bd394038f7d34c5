mkdir -p gpurun_out
timeout 900 python tests/gpu_fused2_ablate.py 2>&1 | tee gpurun_out/r2_fused2_ablate.txt
timeout 1500 python -m pytest tests -q -x -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log
