mkdir -p gpurun_out
SECONDS=0
timeout 900 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo "bench exit $? after $SECONDS s"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_default.json').read().strip().splitlines()[-1])
print('ms_per_step', d['ms_per_step'], 'frac', d['roofline']['frac'], 'traffic', d['roofline']['traffic'], 'launches', d['gpu_launches'])
print('e2e', d['e2e']); print('cpu', d['cpu_baseline']); print('clocks', d['clocks'])
for k,v in d['other_configs'].items():
    print(k, v if isinstance(v,str) else (round(v['ms_per_pass'],3), {kk: (round(vv,4) if isinstance(vv,float) else vv) for kk,vv in v['roofline'].items() if kk in ('frac','frac_issued','frac_useful','achieved')}))
PY
tail -3 gpurun_out/r2_bench_default.err
SECONDS=0
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 | cut -c1-400; echo "reference arm: $SECONDS s"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
