mkdir -p gpurun_out
NP=${NP:-8}
nvidia-smi topo -m 2>&1 | head -12
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29531 tests/multi_gpu_check.py > gpurun_out/r2_multi_n$NP.log 2>&1; echo "multi exit $?"; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/r2_multi_n$NP.log | tail -16
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $NP --steps 100 --warmup 5 > gpurun_out/r2_bench_n$NP.json 2> gpurun_out/r2_bench_n$NP.err; echo "bench exit $?"; python -c "import sys,json; d=json.loads(open('gpurun_out/r2_bench_n$NP.json').read().strip().splitlines()[-1]); print('strong', d['ms_per_step'], 'weak', d['weak']['ms_per_step'], 'e2e', d['e2e'], d['clocks'])"; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/r2_bench_n$NP.err | tail -3
timeout 300 python bench.py --steps 100 --warmup 5 --no-e2e --no-cpu-baseline --no-other-configs 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=1', d['ms_per_step'], d['roofline']['frac'], d['clocks'])"
