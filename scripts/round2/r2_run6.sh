mkdir -p gpurun_out
NP=${NP:-8}
timeout 300 python -m pytest tests/test_gpu_peer.py -q -x -m gpu 2>&1 | tail -3
export BB_LIB_PATH=$PWD/bayesic_b200/lib/libbayesic_b200_timeline.so
echo "--- timeline $NP ranks, 2 Mi rows each"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29541 tests/gpu_timeline.py 2097152 2>&1 | grep "last CTA" | tail -6
unset BB_LIB_PATH
echo "--- bench N=$NP"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $NP --steps 100 --warmup 5 > gpurun_out/r2_bench_n$NP.json 2> gpurun_out/r2_bench_n$NP.err; echo "bench exit $?"; python -c "import sys,json; d=json.loads(open('gpurun_out/r2_bench_n$NP.json').read().strip().splitlines()[-1]); print('strong', d['ms_per_step'], 'weak', d['weak']['ms_per_step'], 'e2e', d['e2e'], d['impl_notes']['collective'][:60])"; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/r2_bench_n$NP.err | tail -4
