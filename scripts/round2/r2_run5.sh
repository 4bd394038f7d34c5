mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_peer.py tests/test_gpu_stats.py tests/test_gpu_passes.py -q -x -m gpu 2>&1 | tail -8
export BB_LIB_PATH=$PWD/bayesic_b200/lib/libbayesic_b200_timeline.so
echo "--- timeline 1 rank, 18944 / 2 Mi rows"
timeout 120 python tests/gpu_timeline.py 18944 2>&1 | tail -4
timeout 120 python tests/gpu_timeline.py 2097152 2>&1 | tail -4
echo "--- timeline 2 ranks, 2 Mi rows each (rank 0 and rank 1 both print)"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/gpu_timeline.py 2097152 2>&1 | grep -v "OMP_NUM\|\*\*\*\*" | tail -12
unset BB_LIB_PATH
echo "--- bench N=1 (16 Mi, then 2 Mi rows)"
timeout 300 python bench.py --steps 50 --warmup 5 --no-e2e --no-cpu-baseline --no-other-configs 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['rows_total'], d['ms_per_step'], d['roofline']['frac'])"
timeout 300 python bench.py --steps 200 --warmup 5 --rows-total 2097152 --no-e2e --no-cpu-baseline --no-other-configs 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['rows_total'], d['ms_per_step'], d['roofline']['frac'])"
echo "--- bench N=2"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 50 --warmup 3 --no-e2e > gpurun_out/r2_bench_n2b.json 2> gpurun_out/r2_bench_n2b.err; echo "bench exit $?"; python -c "import sys,json; d=json.loads(open('gpurun_out/r2_bench_n2b.json').read().strip().splitlines()[-1]); print('strong', d['ms_per_step'], 'weak', d['weak']['ms_per_step'], d['impl_notes']['collective'][:60])"; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/r2_bench_n2b.err | tail -4
