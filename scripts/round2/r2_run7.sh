mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_peer.py tests/test_gpu_stats.py -q -x -m gpu 2>&1 | tail -3
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_peer.py -q -x -m gpu -k "gaussian_pass" 2>&1 | tail -5
export BB_LIB_PATH=$PWD/bayesic_b200/lib/libbayesic_b200_timeline.so
timeout 120 python tests/gpu_timeline.py 2097152 2>&1 | tail -2
unset BB_LIB_PATH
timeout 300 python bench.py --steps 200 --warmup 5 --rows-total 2097152 --no-e2e --no-cpu-baseline --no-other-configs 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['rows_total'], d['ms_per_step'], d['roofline']['frac'])"
