mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py --steps 50 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"; tail -3 gpurun_out/bench_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print('value',d['value'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'])
for k,v in d.get('other_configs',{}).items(): print(k, v if isinstance(v,str) else (round(v['ms_per_pass'],3), round(v['points_per_s']/1e6,1), v['roofline']['bound'], round(v['roofline']['frac'],3)))
PY
