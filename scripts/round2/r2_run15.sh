mkdir -p gpurun_out
NP=${NP:-8}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29551 tests/gpu_h2d_probe.py 512 2>&1 | grep -v "OMP_NUM\|\*\*\*\*" | tail -3
timeout 120 python tests/gpu_h2d_probe.py 512 2>&1 | tail -1
export BB_LIB_PATH=$PWD/bayesic_b200/lib/libbayesic_b200_timeline.so
export BB_SUFFSTATS_PDL=0
echo "--- timeline $NP ranks, 2 Mi rows each (plain launches so that the stamps of consecutive launches do not interleave)"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29541 tests/gpu_timeline.py 2097152 2>&1 | grep "last CTA" | tail -8
