mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_peer.py -q -x -m gpu 2>&1 | tail -15
for rows in 18944 2097152; do
timeout 300 python bench.py --steps 200 --warmup 5 --rows-total $rows --graph 1 --no-e2e --no-cpu-baseline --no-other-configs 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['rows_total'], 'graph', d['ms_per_step'])"
timeout 300 python bench.py --steps 200 --warmup 5 --rows-total $rows --no-e2e --no-cpu-baseline --no-other-configs 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['rows_total'], 'stream', d['ms_per_step'])"
done
timeout 1500 python -m pytest tests -q -x -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -c 900
