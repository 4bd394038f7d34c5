mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv | tail -2
timeout 600 python -m pytest tests/test_gpu_peer.py -q -x -m gpu 2>&1 | tail -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tests/multi_gpu_check.py > gpurun_out/r2_multi_n2.log 2>&1; echo "multi exit $?"; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/r2_multi_n2.log | tail -30
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 50 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench exit $?"; cat gpurun_out/r2_bench_n2.json; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/r2_bench_n2.err | tail -4
