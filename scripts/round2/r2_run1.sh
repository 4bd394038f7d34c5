mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv | tail -1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_peer.py tests/test_gpu_streaming.py -q -x -m gpu 2>&1 | tail -15
timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --no-other-configs 2>&1 | tail -c 1500
timeout 300 python bench.py --steps 30 --warmup 5 --graph 1 --no-e2e --no-cpu-baseline --no-other-configs 2>&1 | tail -c 1500
timeout 300 python bench.py --steps 30 --warmup 5 --rows-total 2097152 --no-e2e --no-cpu-baseline --no-other-configs 2>&1 | tail -c 600
timeout 300 python bench.py --steps 30 --warmup 5 --rows-total 2097152 --graph 1 --no-e2e --no-cpu-baseline --no-other-configs 2>&1 | tail -c 600
timeout 1500 python -m pytest tests -q -x -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/pytest_gpu.log
