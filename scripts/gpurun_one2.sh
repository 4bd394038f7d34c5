mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stats.py tests/test_gpu_passes.py -q -x -m gpu -k "single_entry or cfg2 or gaussian" > gpurun_out/pytest_one.log 2>&1; echo "pytest exit $?"; tail -20 gpurun_out/pytest_one.log
timeout 300 python bench.py --steps 50 --warmup 3 --no-e2e --no-cpu-baseline --no-other-configs 2>gpurun_out/b.err | cut -c1-1400
