mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 50 --warmup 3 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; echo "exit $?"; cut -c1-400 gpurun_out/bench_n4.json; tail -3 gpurun_out/bench_n4.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --impl reference --gpus 4 --steps 2 --warmup 1 > gpurun_out/bench_ref_n4.json 2> gpurun_out/bench_ref_n4.err; echo "ref exit $?"; cut -c1-300 gpurun_out/bench_ref_n4.json
