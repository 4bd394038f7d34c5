mkdir -p gpurun_out
NP=${NP:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29531 tests/multi_gpu_check.py --no-timing 2>&1 | grep -v "OMP_NUM\|\*\*\*\*" | tail -14
for mode in "1 1" "0 0"; do
set -- $mode
BB_SUFFSTATS_DYNAMIC=$1 BB_SUFFSTATS_PDL=$2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $NP --steps 100 --warmup 5 --no-e2e > gpurun_out/r2_bench_n${NP}_$1$2.json 2> gpurun_out/r2_bench_n${NP}_$1$2.err; echo "bench dynamic=$1 pdl=$2 exit $?"; python -c "import sys,json; d=json.loads(open('gpurun_out/r2_bench_n${NP}_$1$2.json').read().strip().splitlines()[-1]); print('strong', d['ms_per_step'], 'weak', d['weak']['ms_per_step'], d['impl_notes']['collective'][:50], d['clocks'])"; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/r2_bench_n${NP}_$1$2.err | tail -3
done
timeout 300 python bench.py --steps 100 --warmup 5 --no-e2e --no-cpu-baseline --no-other-configs 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=1', d['ms_per_step'], d['roofline']['frac'], d['clocks'])"
