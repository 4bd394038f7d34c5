mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stats.py tests/test_gpu_plan_parity.py tests/test_gpu_passes.py -q -x -m gpu 2>&1 | tail -4
timeout 300 python tests/gpu_cfg1_latency.py 2>&1 | tee gpurun_out/r2_cfg1_latency.txt
B="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-other-configs"
timeout 300 $B > gpurun_out/r2_launches_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches.csv $B > gpurun_out/r2_launches_ncu.log 2>&1
tail -1 gpurun_out/r2_launches_plain.log | cut -c1-200; wc -l gpurun_out/r2_launches.csv
ncu --set full --clock-control none --import-source on -k regex:suffstats_tc_kernel -s 3 -c 1 -o gpurun_out/prof_r2_suffstats $B > gpurun_out/r2_ncu_suffstats.log 2>&1; tail -2 gpurun_out/r2_ncu_suffstats.log
ls -la gpurun_out/prof_r2_suffstats.ncu-rep
