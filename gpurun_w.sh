for i in 1 2; do timeout 100 python tests/gpu_profile_driver.py suffstats; done
timeout 300 python bench.py --steps 50 --warmup 3 --no-e2e --no-cpu-baseline --no-other-configs | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'])"
