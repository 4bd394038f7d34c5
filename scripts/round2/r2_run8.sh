mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_peer.py tests/test_gpu_stats.py tests/test_gpu_passes.py -q -x -m gpu 2>&1 | tail -3
b() { timeout 200 python bench.py --steps $1 --warmup 5 --rows-total $2 --no-e2e --no-cpu-baseline --no-other-configs 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$3', d['config']['rows_total'], round(d['ms_per_step']*1000,2), 'us  frac', round(d['roofline']['frac'],4), 'elbo', d['elbo'])"; }
for rows in 2097152 16777216; do
  BB_SUFFSTATS_DYNAMIC=0 BB_SUFFSTATS_PDL=0 b 200 $rows "static/plain  "
  BB_SUFFSTATS_DYNAMIC=1 BB_SUFFSTATS_PDL=0 b 200 $rows "dynamic/plain "
  BB_SUFFSTATS_DYNAMIC=0 BB_SUFFSTATS_PDL=1 b 200 $rows "static/pdl    "
  BB_SUFFSTATS_DYNAMIC=1 BB_SUFFSTATS_PDL=1 b 200 $rows "dynamic/pdl   "
done
export BB_LIB_PATH=$PWD/bayesic_b200/lib/libbayesic_b200_timeline.so
timeout 120 python tests/gpu_timeline.py 2097152 2>&1 | tail -4
