mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_peer.py -q -x -m gpu 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tests/multi_gpu_check.py --no-timing 2>&1 | grep -v "OMP_NUM\|\*\*\*\*" | tail -12
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/r2_bench_n2_final.json 2> gpurun_out/r2_bench_n2_final.err; echo "bench exit $?"; python -c "import sys,json; d=json.loads(open('gpurun_out/r2_bench_n2_final.json').read().strip().splitlines()[-1]); print('strong', d['ms_per_step'], 'weak', d['weak']['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['clocks'])"
