mkdir -p gpurun_out
timeout 420 python -m pytest tests -q -x -m gpu 2>&1 | tail -3
timeout 200 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_final.json').read().strip().splitlines()[-1])
print('value', d['value'], d['unit'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'])
for k,v in d.get('other_configs',{}).items():
    print(k, {kk: v[kk] for kk in v if kk in ('ms_per_step','value','unit')}, v.get('roofline',{}).get('frac', v.get('roofline',{}).get('frac_issued')))
print(d['clocks'])
PY
for rep in 1 2; do timeout 120 python tests/gpu_profile_driver.py gram 2>&1 | tail -1; done
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
